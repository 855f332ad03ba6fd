#!/bin/bash
# 2-GPU pass: partition tests (p2p + nccl), partitioned bench with both exchanges
mkdir -p gpurun_out/r02b
cd /root/repo
nvidia-smi topo -m > gpurun_out/r02b/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_partition.py -x -q -s > gpurun_out/r02b/pytest_partition.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/pytest_partition.log
tail -15 gpurun_out/r02b/pytest_partition.log
for x in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --exchange $x > gpurun_out/r02b/bench_n2_$x.json 2> gpurun_out/r02b/bench_n2_$x.err; echo "bench $x rc=$?"
tail -c 1500 gpurun_out/r02b/bench_n2_$x.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b/bench_n2_full.json 2> gpurun_out/r02b/bench_n2_full.err; echo "bench full rc=$?"
tail -c 2500 gpurun_out/r02b/bench_n2_full.json
tail -5 gpurun_out/r02b/bench_n2_full.err
