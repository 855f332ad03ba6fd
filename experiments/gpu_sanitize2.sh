#!/bin/bash
# compute-sanitizer over the n = 16 paths of the final build (coalesced back substitution, cluster segment kernel)
mkdir -p gpurun_out/r02final
cd /root/repo
SEL="headline or chain or test_update_is_independent or dichotomic"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r02final/sanitizer_memcheck_n16.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02final/sanitizer_memcheck_n16.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "headline" > gpurun_out/r02final/sanitizer_racecheck_n16.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02final/sanitizer_racecheck_n16.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "headline" > gpurun_out/r02final/sanitizer_synccheck_n16.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r02final/sanitizer_synccheck_n16.log
