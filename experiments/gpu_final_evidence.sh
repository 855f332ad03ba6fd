#!/bin/bash
# final evidence of the round on the final build: GPU suite, default bench line, launch list, ncu --set full of every kernel
# of one C2 Newton step (-> profiles/r02_traffic.json) and of the adaptive-path kernels (defect, mesh selection, interpolation)
out=gpurun_out/r02final
mkdir -p $out
cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
# one step under --set full: skip the warm-up launches, capture 3 steps' worth (9 kernels each); the last complete step is used
# (the .ncu-rep stays on the box: gpurun_out is capped at 64 MiB; the summaries are made here)
timeout 1500 ncu --set full --clock-control none -k regex:'k_resjac|k_bc|k_reduce|k_tail|k_seg_cluster|k_backsub' -c 36 -o /tmp/c2_step -f python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
SHA=$(python -c "import bench; print(bench.kernel_source_sha16())")
python profiles/make_traffic.py /tmp/c2_step.ncu-rep $SHA 20000 > $out/r02_traffic.json; echo "traffic rc=$?"
python profiles/summarize_ncu.py full /tmp/c2_step.ncu-rep | python -c "
import sys
# keep the LAST step's launches (9 kernels): split on '---'
blocks = sys.stdin.read().split('---')
print(blocks[0].rstrip()); print('---' + '---'.join(blocks[-9:]))" > $out/full_c2_step.txt
for k in k_resjac_tape k_reduce_warp k_seg_cluster16 k_backsub_warp; do python profiles/stall_breakdown.py /tmp/c2_step.ncu-rep $k 2 >> $out/stalls_c2_step.txt 2>/dev/null; done; echo "stalls rc=$?"
# source-level capture of the dominant kernel only (small enough to travel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_warp -s 2 -c 1 -o $out/l0_reduce -f python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_l0.log 2>&1; echo "ncu l0 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'k_defect|k_mesh_select|k_interp|k_half_mesh' -c 8 -o /tmp/adaptive -f python -c "
import math, mirk_b200 as M
from boundaryvaluediffeq_jl_b200 import configs
c = configs.c2_chain8(1999)
sol = M.solve(M.BVProblem(c.problem, c.y0, c.tspan, p=c.p, mesh=c.mesh), M.MIRK4(max_num_subintervals=20000), adaptive=True, abstol=1e-9)
print(sol.retcode, len(sol.t), sol.original['hist_n_mesh'])
" > $out/ncu_adaptive.log 2>&1; echo "ncu adaptive rc=$?"; tail -2 $out/ncu_adaptive.log
python profiles/summarize_ncu.py full /tmp/adaptive.ncu-rep > $out/full_adaptive.txt
ls -la $out
