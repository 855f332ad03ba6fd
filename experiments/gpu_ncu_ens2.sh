#!/bin/bash
# source-level profile of the warp-per-trajectory ensemble kernel (65 536 trajectories): hot source lines
mkdir -p gpurun_out/r02final
cd /root/repo
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ensemble_warp -s 1 -c 1 -o /tmp/ens_warp -f python bench.py --workload c3 --steps 1 --trajectories 65536 > gpurun_out/r02final/ncu_ens.log 2>&1; echo "ncu rc=$?"
python profiles/source_hotspots.py /tmp/ens_warp.ncu-rep k_ensemble_warp 0 60 > gpurun_out/r02final/ens_hotspots.txt 2>&1; head -70 gpurun_out/r02final/ens_hotspots.txt
python profiles/stall_breakdown.py /tmp/ens_warp.ncu-rep k_ensemble_warp 0 > gpurun_out/r02final/ens_stalls.txt 2>&1; cat gpurun_out/r02final/ens_stalls.txt | head -30
