#!/bin/bash
# compute-sanitizer memcheck over the paths the earlier runs did not touch: large blocks (n = 32 four-warp merge, n = 128 blocked
# Gauss-Jordan, stage-wise Jacobian), nonlinear-solver fallbacks, global-error controllers, the standalone ABD solver, MIRK6I
mkdir -p gpurun_out/r02v
cd /root/repo
SEL="large_block_problems or nonlinear_solvers or global_error or polyalgorithm or standalone_abd or failure_paths"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r02v/sanitizer_memcheck_wide.log 2>&1; echo "memcheck wide rc=$?"; tail -4 gpurun_out/r02v/sanitizer_memcheck_wide.log
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zz_mirk6i.py -m gpu -x -q > gpurun_out/r02v/sanitizer_memcheck_6i.log 2>&1; echo "memcheck 6i rc=$?"; tail -3 gpurun_out/r02v/sanitizer_memcheck_6i.log
