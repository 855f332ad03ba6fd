#!/bin/bash
# sweep of the reduction shape (level-0 chunk c0, upper chunk c1, tail threshold) through bench.py --chunk
out=gpurun_out/s3/sweep_chunk.txt
: > $out
for cfg in "12 4 16" "8 4 16" "6 4 16" "4 4 16" "16 4 16" "12 2 16" "12 3 16" "12 8 16" "6 3 16" "4 2 16" "8 2 16" "12 4 8" "12 4 32" "12 4 64" "6 2 32"; do
  set -- $cfg
  ch=$(( $1 | ($2 << 8) | ($3 << 16) ))
  line=$(python bench.py --steps 20 --warmup 3 --chunk $ch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get('phases_ms_per_step',{})
print('%.4f ms  ' % d['ms_per_step'] + ' '.join('%s=%.1f' % (k.split('+')[0][:14], 1e3*v) for k,v in p.items()))")
  echo "c0=$1 c1=$2 tail=$3 : $line" | tee -a $out
done
