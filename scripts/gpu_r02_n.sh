#!/bin/bash
# round 2, call n: fused GCN/GIN as opt-in — kernel + layer tests, host pipeline, partition-virtual tests, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gcn_fused.py tests/test_gpu_segsum.py tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r02n_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02n_tests.log
tail -5 gpurun_out/r02n_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02n_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline'], d['e2e'], d.get('parity_check'))
PY
