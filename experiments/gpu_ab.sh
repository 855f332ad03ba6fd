#!/bin/bash
# A/B of an environment switch on the default bench line: gpu_ab.sh "VAR=0" "VAR=1" ... (after the n = 16 parity tests with the last setting)
mkdir -p gpurun_out/r02s2
cd /root/repo
for v in "$@"; do
  echo "== $v"
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain or headline or abd or c2" 2>&1 | tail -2
  env $v timeout 300 python bench.py --no-extra --steps 20 --warmup 3 2> gpurun_out/r02s2/bench_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.4f  e2e %.1f  launches %s' % (d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print({k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})"
  tail -2 gpurun_out/r02s2/bench_ab.err
done
