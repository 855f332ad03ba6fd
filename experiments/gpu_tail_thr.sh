#!/bin/bash
# tail threshold (relations from which the one-block tail kernel takes over) with the cluster segment kernel in place
cd /root/repo
for thr in 2 4 8 16 32 64; do
python bench.py --no-extra --steps 20 --warmup 3 --chunk $((thr << 16)) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tail_thr $thr: ms/step %.4f' % d['ms_per_step'], {k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items() if 'abd' in k})"
done
