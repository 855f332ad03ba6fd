#!/bin/bash
# 2-GPU pass on HEAD: the mesh-partitioned GPU test (2 ranks, both transports) and the driver's bench line at N = 2
mkdir -p gpurun_out/r02v
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_partition.py -m gpu -x -q -s > gpurun_out/r02v/pytest_partition_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02v/pytest_partition_2gpu.log
bash experiments/gpu_run8.sh 2
cp gpurun_out/r02h/bench_n2.json gpurun_out/r02v/bench_n2.json
