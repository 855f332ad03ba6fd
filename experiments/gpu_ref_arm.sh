#!/bin/bash
# the reference arm (CPU restatement on the GPU box's host) and the default bench line beside it
mkdir -p gpurun_out/r02v
cd /root/repo
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v/bench_ref.json 2> gpurun_out/r02v/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r02v/bench_n1_final.json 2> gpurun_out/r02v/bench_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r02v/bench_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r02v/bench_n1_final.json').read().strip().splitlines()[-1])
print('reference arm: %.3f steps/s (%.1f ms/step), %s threads; one thread %.3f' % (r['value'], r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_baseline']['single_thread_value']))
print('b200: value %.1f e2e %.1f cpu_baseline %.3f (%s threads; 1 thread %.3f)' % (d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['single_thread_value']))
print('c3 cpu', d['extra']['c3_strong']['cpu_baseline']['value'], 'c1', {k: round(v['cpu_port_ms'],3) for k,v in d['extra']['c1']['algs'].items()})
PY
