#!/bin/bash
mkdir -p gpurun_out/r02c
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bratu64 or large_block or full_size" > gpurun_out/r02c/pytest_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c/pytest_block.log
tail -15 gpurun_out/r02c/pytest_block.log
timeout 300 python bench.py --workload c4 --steps 5 > gpurun_out/r02c/c4_dense.json 2>&1; tail -c 1200 gpurun_out/r02c/c4_dense.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c/launches_c4.csv python bench.py --workload c4 --steps 1 > gpurun_out/r02c/ncu_c4.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02c/launches_c4.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<15: continue
    k=r[4].split('(')[0][:60]+" grid="+r[8]
    agg[k][0]+=1; agg[k][1]+=float(r[14])/1e3
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f"{t:10.1f} us  x{c:3d}  {k}")
PY
