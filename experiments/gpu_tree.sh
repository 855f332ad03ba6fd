#!/bin/bash
# tree kernels (abd_team.cuh): parity tests on the n = 16 paths, then A/B bench lines (MIRK_TREE=0 is the old segment/tail path)
mkdir -p gpurun_out/r02s2
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${1:-chain or headline or abd or large or c2}" > gpurun_out/r02s2/pytest_tree.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02s2/pytest_tree.log
for v in "MIRK_TREE=0" "MIRK_TREE=1" "MIRK_TREE=1 MIRK_TREE_TEAM=2" "MIRK_TREE=1 MIRK_TREE_TEAM=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --no-extra --steps 20 --warmup 3 2> gpurun_out/r02s2/bench_tree.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.4f  e2e %.1f  launches %s' % (d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print({k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})"
  tail -2 gpurun_out/r02s2/bench_tree.err
done
