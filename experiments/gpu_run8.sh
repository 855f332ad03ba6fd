#!/bin/bash
# multi-GPU pass: the driver's bench line at N GPUs (partitioned C2 + extras)
N=${1:-8}
mkdir -p gpurun_out/r02h
cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02h/bench_n$N.json 2> gpurun_out/r02h/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02h/bench_n$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_vs_single_gpu']['max_rel_diff'] if d['parity_vs_single_gpu'] else None)
print('phases', {k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()})
x=d['extra']
print('c3', x['c3_strong'].get('value'), x['c3_strong'].get('ms_per_step'), x['c3_strong'].get('converged_fraction'), x['c3_strong'].get('error'))
print('c5', x['c5part'].get('ms_per_step'), x['c5part'].get('phases_ms_rank0'), x['c5part'].get('error'))
PY
tail -3 gpurun_out/r02h/bench_n$N.err
