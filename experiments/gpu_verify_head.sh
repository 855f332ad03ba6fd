#!/bin/bash
# verification of HEAD on one GPU: the whole GPU suite, smoke, and the default bench line with its wall time
out=gpurun_out/r02v
mkdir -p $out
cd /root/repo
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
t1=$(date +%s); echo "pytest wall $((t1-t0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
t2=$(date +%s)
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
t3=$(date +%s); echo "bench wall $((t3-t2)) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step %.4f value %.1f e2e %.1f traffic %s frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['traffic'], d['roofline']['frac']))
print({k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})
e=d.get('extra',{})
for k,v in e.items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
