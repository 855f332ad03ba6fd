#!/bin/bash
mkdir -p gpurun_out/r02e
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_ensemble.py -x -q > gpurun_out/r02e/pytest_ens.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e/pytest_ens.log
tail -25 gpurun_out/r02e/pytest_ens.log
for nc in 64 96 128; do
MIRK_ENS_SMEM_NODES=$nc timeout 300 python bench.py --workload c3 --steps 3 > gpurun_out/r02e/c3_warp_$nc.json 2>&1; echo "NCs=$nc"; tail -c 900 gpurun_out/r02e/c3_warp_$nc.json | cut -c1-700
done
MIRK_ENS_KERNEL=thread timeout 300 python bench.py --workload c3 --steps 3 > gpurun_out/r02e/c3_thread.json 2>&1; tail -c 900 gpurun_out/r02e/c3_thread.json | cut -c1-400
