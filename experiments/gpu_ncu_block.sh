#!/bin/bash
mkdir -p gpurun_out/r02d
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_reduce_block -s 0 -c 1 -o gpurun_out/r02d/block_l0 -f python bench.py --workload c4 --steps 1 > gpurun_out/r02d/ncu_block.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_gemm -s 0 -c 1 -o gpurun_out/r02d/chain_gemm -f python bench.py --workload c4 --steps 1 > gpurun_out/r02d/ncu_gemm.log 2>&1
ls -la gpurun_out/r02d
