#!/bin/bash
# multi-GPU pass: partition tests (p2p + nccl), partitioned bench
N=${1:-2}
mkdir -p gpurun_out/r02b
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_partition.py -x -q -s > gpurun_out/r02b/pytest_partition_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/pytest_partition_n$N.log
tail -14 gpurun_out/r02b/pytest_partition_n$N.log
for x in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --exchange $x > gpurun_out/r02b/bench_n${N}_$x.json 2> gpurun_out/r02b/bench_n${N}_$x.err; echo "bench $x rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02b/bench_n${N}_$x.json').read().strip().splitlines()[-1]); print('$x', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k: round(v*1e3,1) for k,v in d['phases_ms_per_step'].items()}, d['parity_vs_single_gpu']['max_rel_diff'], d['gpu_launches'])"
done
MIRK_PART_FUSED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --exchange p2p > gpurun_out/r02b/bench_n${N}_p2p_unfused.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02b/bench_n${N}_p2p_unfused.json').read().strip().splitlines()[-1]); print('unfused', d['value'], d['ms_per_step'])"
