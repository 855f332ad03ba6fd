#!/bin/bash
# compute-sanitizer over the ensemble kernels and the mesh-partitioned path (one rank: peer-memory push / wait kernels, interface solve)
mkdir -p gpurun_out/r02final
cd /root/repo
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q > gpurun_out/r02final/sanitizer_memcheck_ens.log 2>&1; echo "memcheck ensemble rc=$?"; tail -3 gpurun_out/r02final/sanitizer_memcheck_ens.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q -k "matches or oracle" > gpurun_out/r02final/sanitizer_racecheck_ens.log 2>&1; echo "racecheck ensemble rc=$?"; tail -3 gpurun_out/r02final/sanitizer_racecheck_ens.log
MASTER_ADDR=127.0.0.1 MASTER_PORT=29544 RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/partition_worker.py > gpurun_out/r02final/sanitizer_memcheck_part.log 2>&1; echo "memcheck partition rc=$?"; grep -v "^W1017\|^\[W" gpurun_out/r02final/sanitizer_memcheck_part.log | tail -4
