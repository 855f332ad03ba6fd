#!/bin/bash
# source-level capture of the four k_tail_warp launches of one C2 step (upper segment, one-block tail, two back-substitution segments)
out=gpurun_out/r02t
mkdir -p $out
cd /root/repo
timeout 600 ncu --section SourceCounters --section WarpStateStats --section LaunchStats --section SpeedOfLight --clock-control none --import-source on -k regex:k_tail_warp -s 8 -c 4 -o $out/tail -f python bench.py --steps 2 --warmup 1 --no-extra --profile > $out/ncu_tail.log 2>&1; echo "ncu rc=$?"
tail -3 $out/ncu_tail.log
ls -la $out
