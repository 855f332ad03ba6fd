#!/bin/bash
# round 2, first GPU pass: full GPU test-suite, bench line with extras, launch list
mkdir -p gpurun_out/r02a
cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest_gpu.log
tail -5 gpurun_out/r02a/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a/bench_ref.json 2>&1
MIRK_ABD_MMA32=0 timeout 300 python bench.py --workload c5part --c5-nint 249999 --steps 5 > gpurun_out/r02a/c5_250k_pair.json 2>&1
timeout 300 python bench.py --workload c5part --c5-nint 249999 --steps 5 > gpurun_out/r02a/c5_250k_mma32.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a/launches.csv python bench.py --steps 2 --warmup 3 --profile --no-extra > gpurun_out/r02a/ncu_bench.log 2>&1
echo done
