#!/bin/bash
# round 2, call w: the cfg5-shaped train step of the default N = 8 line on ONE GPU's share (12.5 M cells): time and peak memory
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --train-step --train-checkpoint --train-cells 12500000 > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d.get('train_step_partitioned'))
PY
tail -3 gpurun_out/r02w_bench.err
