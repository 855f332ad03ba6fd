#!/bin/bash
mkdir -p gpurun_out/r02g
cd /root/repo
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r02g/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g/pytest_gpu.log
tail -60 gpurun_out/r02g/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02g/bench.json 2> gpurun_out/r02g/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02g/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases_ms_per_step'])"
