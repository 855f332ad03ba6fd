#!/bin/bash
mkdir -p gpurun_out/r02f
cd /root/repo
MIRK_ENS_SMEM_NODES=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ensemble_warp -s 1 -c 1 -o gpurun_out/r02f/ens_warp -f python bench.py --workload c3 --steps 1 --trajectories 65536 > gpurun_out/r02f/ncu_ens.log 2>&1
ls -la gpurun_out/r02f
