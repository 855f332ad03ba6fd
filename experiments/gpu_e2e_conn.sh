#!/bin/bash
# end-to-end leg: hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) x pipelined handles
cd /root/repo
for conn in 8 32; do
  for h in 12 16 24; do
    CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python bench.py --no-extra --steps 20 --warmup 3 --e2e-handles $h 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('conn $conn handles $h: ms/step %.4f  e2e %.1f  conc %.1f (%s handles)' % (d['ms_per_step'], d['e2e']['value'], d['concurrent_problems']['value'], d['concurrent_problems']['handles']))"
  done
done
