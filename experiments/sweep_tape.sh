#!/bin/bash
# sweep of the taped Jacobian kernel's shape: warps per CTA x intervals per CTA (MIRK_TAPE_WARPS / MIRK_TAPE_IPW)
out=$1; shift
: > $out
for cfg in "$@"; do
  set -- $cfg
  line=$(MIRK_TAPE_WARPS=$1 MIRK_TAPE_IPW=$2 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get('phases_ms_per_step',{})
print('%.4f ms  resjac=%.1f us' % (d['ms_per_step'], 1e3*p['residual+jacobian_blocks']))")
  echo "warps=$1 ipw=$2 : $line" | tee -a $out
done
