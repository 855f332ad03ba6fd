#!/bin/bash
# the default bench line without the extras: checks that bench.py runs and prints the roofline block
mkdir -p gpurun_out/r02v
cd /root/repo
timeout 200 python bench.py --no-extra > gpurun_out/r02v/bench_noextra.json 2> gpurun_out/r02v/bench_noextra.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v/bench_noextra.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mode"][:90])
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
tail -2 gpurun_out/r02v/bench_noextra.err
