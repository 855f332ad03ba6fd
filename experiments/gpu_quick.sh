#!/bin/bash
# quick check of a kernel change: n = 16 parity tests, then the default bench line without extras
mkdir -p gpurun_out/r02s2
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${1:-chain or headline or abd or large or c2}" > gpurun_out/r02s2/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02s2/pytest_quick.log
timeout 300 python bench.py --no-extra --steps 20 --warmup 3 2> gpurun_out/r02s2/bench_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.4f  e2e %.1f  launches %s' % (d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print({k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})"
tail -2 gpurun_out/r02s2/bench_quick.err
