#!/bin/bash
# launch list (ncu, durations only) of `bench.py --profile` under the environment given as arguments: gpu_launchlist.sh tag VAR=val ...
tag=$1; shift
mkdir -p gpurun_out/r02s2
cd /root/repo
env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02s2/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-extra --profile > gpurun_out/r02s2/ncu_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02s2/launches_$tag.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]: print(r[4][:70], r[7], r[8], r[-1])
PY
