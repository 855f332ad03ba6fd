#!/bin/bash
# C4 (n = 128): panel width of the blocked Gauss-Jordan (experiments/bin/libmirkb200_nb{32,16,8}.so = builds with -DMIRK_BLOCK_NB)
cd /root/repo
cp boundaryvaluediffeq.jl_b200/libmirkb200.so /tmp/lib_orig.so
for nb in 32 16 8; do
  cp experiments/bin/libmirkb200_nb$nb.so boundaryvaluediffeq.jl_b200/libmirkb200.so
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_block or bratu64" 2>&1 | tail -1
  python bench.py --workload c4 --steps 8 2>/dev/null > /dev/null
  python bench.py --workload c4 --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NB $nb: ms/step %.2f' % d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_rank0'].items() if v > 0.05})"
done
cp /tmp/lib_orig.so boundaryvaluediffeq.jl_b200/libmirkb200.so
