#!/bin/bash
# correctness evidence of the session: the new full-size C5 test, then compute-sanitizer (memcheck, racecheck) over a
# subset of the GPU suite that exercises every ABD reduction path (n = 2..8 warp, n = 16 DMMA, n = 32, n = 128 block)
mkdir -p gpurun_out/r02s2
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c5_full_size" > gpurun_out/r02s2/pytest_c5full.log 2>&1; echo "c5 full rc=$?"; tail -3 gpurun_out/r02s2/pytest_c5full.log
SEL="test_residual_jacobian_and_update_match_oracle or test_large_block_problems_match_golden or test_standalone_abd or test_defect_and_mesh"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r02s2/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02s2/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_large_block_problems_match_golden or test_standalone_abd or chain8" > gpurun_out/r02s2/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02s2/sanitizer_racecheck.log
