#!/bin/bash
# A/B of alternative builds of the library: gpu_ab_la.sh lib1.so lib2.so ...  ("-" = the in-tree build); the parity tests of
# the n = 16 paths run on every alternative build
cd /root/repo
L=boundaryvaluediffeq.jl_b200/libmirkb200.so
cp $L /tmp/orig.so
for lib in "$@"; do
  if [ "$lib" = "-" ]; then cp /tmp/orig.so $L; else cp $lib $L; fi
  echo "== $lib"
  if [ "$lib" != "-" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain or headline or abd or c2 or large" 2>&1 | tail -2
  fi
  for rep in 1 2; do
    timeout 300 python bench.py --no-extra --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.4f  e2e %.1f  conc %.1f' % (d['ms_per_step'], d['e2e']['value'], d['concurrent_problems']['value']))
print({k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})"
  done
done
cp /tmp/orig.so $L
