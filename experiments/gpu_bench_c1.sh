#!/bin/bash
# default bench line with every extra (c3, c5, c4, c1) and its wall time
mkdir -p gpurun_out/r02v
cd /root/repo
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02v/bench_n1_c1.json 2> gpurun_out/r02v/bench_n1_c1.err; echo "bench rc=$?"
t1=$(date +%s); echo "bench wall $((t1-t0)) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v/bench_n1_c1.json').read().strip().splitlines()[-1])
print('ms/step %.4f value %.1f e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
print(json.dumps(d['extra'].get('c1'), indent=1))
PY
tail -3 gpurun_out/r02v/bench_n1_c1.err
