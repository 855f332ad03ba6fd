#!/bin/bash
# racecheck / synccheck over the large-block kernels (n = 32 four-warp merge, n = 64 / 128 blocked Gauss-Jordan, stage-wise Jacobian)
mkdir -p gpurun_out/r02v
cd /root/repo
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_block_problems" > gpurun_out/r02v/sanitizer_racecheck_large.log 2>&1; echo "racecheck large rc=$?"; tail -4 gpurun_out/r02v/sanitizer_racecheck_large.log
timeout 120 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_block_problems" > gpurun_out/r02v/sanitizer_synccheck_large.log 2>&1; echo "synccheck large rc=$?"; tail -3 gpurun_out/r02v/sanitizer_synccheck_large.log
