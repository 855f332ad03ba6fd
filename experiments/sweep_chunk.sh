#!/bin/bash
# sweep of the reduction shape (level-0 chunk c0, upper chunk c1 [0 = multi-level radix-2 segments], tail threshold)
# through bench.py --chunk;  usage: sweep_chunk.sh out.txt "c0 c1 tail" ...
out=$1; shift
: > $out
for cfg in "$@"; do
  set -- $cfg
  ch=$(( $1 | ($2 << 8) | ($3 << 16) ))
  line=$(python bench.py --steps 20 --warmup 3 --chunk $ch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get('phases_ms_per_step',{})
print('%.4f ms  launches/step=%.1f  ' % (d['ms_per_step'], d['gpu_launches']/d['steps']) + ' '.join('%s=%.1f' % (k.split('+')[0][:14], 1e3*v) for k,v in p.items()))")
  echo "c0=$1 c1=$2 tail=$3 : $line" | tee -a $out
done
