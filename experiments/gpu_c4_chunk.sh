#!/bin/bash
# C4 (n = 128): level-0 group size against the L2 footprint of the per-CTA working slabs (768 KB each)
cd /root/repo
for c in 0 32 40 48 56 64; do
python bench.py --workload c4 --chunk $c --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $c: ms/step %.2f' % d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_rank0'].items() if v > 0.05})"
done
