#!/bin/bash
mkdir -p gpurun_out/r02c
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "large_block or full_size or abd" > gpurun_out/r02c/pytest_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c/pytest_block.log
tail -15 gpurun_out/r02c/pytest_block.log
timeout 300 python bench.py --workload c4 --steps 5 > gpurun_out/r02c/c4_block.json 2>&1; tail -c 1200 gpurun_out/r02c/c4_block.json
MIRK_ABD_BLOCK=0 timeout 300 python bench.py --workload c4 --steps 2 > gpurun_out/r02c/c4_generic.json 2>&1; tail -c 800 gpurun_out/r02c/c4_generic.json
