#!/bin/bash
# ensemble kernels: GPU tests, then the C3 line
cd /root/repo
python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c3 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c3: %.3f M solves/s, %.2f ms, converged %.4f, newton %.3f, nodes %.2f, e2e %.3f M/s' % (d['value']/1e6, d['ms_per_step'], d['converged_fraction'], d['mean_newton_iters'], d['mean_final_nodes'], d['e2e']['value']/1e6))"
