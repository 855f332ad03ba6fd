#!/bin/bash
# A/B of alternative builds of the library on the GPU box: ab_lib.sh out.txt lib1.so lib2.so ...  ("-" = in-tree build)
out=$1; shift
: > $out
cp boundaryvaluediffeq.jl_b200/libmirkb200.so /tmp/libmirkb200_orig.so
for lib in "$@"; do
  if [ "$lib" = "-" ]; then cp /tmp/libmirkb200_orig.so boundaryvaluediffeq.jl_b200/libmirkb200.so; else cp $lib boundaryvaluediffeq.jl_b200/libmirkb200.so; fi
  for shape in "4 32" "1 8"; do
    set -- $shape
    line=$(MIRK_TAPE_WARPS=$1 MIRK_TAPE_IPW=$2 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get('phases_ms_per_step',{})
print('%.4f ms  e2e %.1f steps/s  resjac=%.1f us' % (d['ms_per_step'], d['e2e']['value'], 1e3*p['residual+jacobian_blocks']))")
    echo "$lib warps=$1 ipw=$2 : $line" | tee -a $out
  done
done
cp /tmp/libmirkb200_orig.so boundaryvaluediffeq.jl_b200/libmirkb200.so
