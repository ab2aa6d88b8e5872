#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02h_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02h_bench.json') if x.startswith('{')]
d=json.loads(l[-1]); ex=d.pop('extras')
print(json.dumps({k:d[k] for k in ('value','ms_per_step','parity_check','e2e','roofline')})[:1500])
for k,v in ex.items(): print(k, json.dumps(v)[:220])
PY
