#!/bin/bash
# final evidence of the round on the final build: GPU suite, default bench line, launch list, ncu --set full of every kernel
# of one C2 Newton step (-> profiles/r02_traffic.json) and of the adaptive-path kernels (defect, mesh selection, interpolation)
out=gpurun_out/r02final
mkdir -p $out
cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
# one step under --set full: skip the warm-up launches, capture 3 steps' worth (9 kernels each); the last complete step is used
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_resjac|k_bc|k_reduce|k_tail|k_backsub' -c 36 -o $out/c2_step -f python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'k_defect|k_mesh_select|k_interp|k_half_mesh' -c 6 -o $out/adaptive -f python -c "
import math, mirk_b200 as M
from boundaryvaluediffeq_jl_b200 import configs
c = configs.c2_chain8(1999)
sol = M.solve(M.BVProblem(c.problem, c.y0, c.tspan, p=c.p, mesh=c.mesh), M.MIRK4(max_num_subintervals=20000), adaptive=True, abstol=1e-9)
print(sol.retcode, len(sol.t), sol.original['hist_n_mesh'])
" > $out/ncu_adaptive.log 2>&1; echo "ncu adaptive rc=$?"; tail -2 $out/ncu_adaptive.log
ls -la $out
