#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 exit $?"; cat gpurun_out/bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','clocks')})
print('e2e',d['e2e']); print('roofline',d['roofline']); print('cpu',d['cpu_baseline'])
for k,v in d['extras'].items(): print(k,v)
"; tail -n 3 gpurun_out/bench_n1.err
