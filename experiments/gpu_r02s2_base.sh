#!/bin/bash
# session-2 baseline of round 2: GPU suite, default bench line, launch list of the same command
mkdir -p gpurun_out/r02s2
cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02s2/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s2/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s2/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02s2/bench_n1.json 2> gpurun_out/r02s2/bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02s2/bench_n1.json | head -c 600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s2/launches.csv python bench.py --steps 2 --warmup 1 --no-extra --profile > gpurun_out/r02s2/ncu_b.log 2>&1; echo "ncu rc=$?"
