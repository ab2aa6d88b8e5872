#!/bin/bash
# round 2, call ai: final single-GPU validation — whole suite, smoke, default bench, cfg5 share of one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02ai_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02ai_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02ai_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02ai_smoke.log
timeout 900 python bench.py > gpurun_out/r02ai_bench.json 2> gpurun_out/r02ai_bench.err; echo "bench exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --train-step --train-checkpoint --train-cells 12500000 > gpurun_out/r02ai_cfg5.json 2> gpurun_out/r02ai_cfg5.err; echo "cfg5 exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ai_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['roofline']['frac'], d['roofline']['kernel_ms']); print(d['e2e']['ms_per_step']); print(d.get('parity_check'))
for k, v in d.get('extras', {}).items():
    if 'bf16_fwd' in k or 'train_step' in k: print(k, {kk: vv for kk, vv in v.items() if kk in ('ms', 'peak_mem_gb')}, (v.get('parity_check') or {}).get('ok'))
c = json.loads(open('gpurun_out/r02ai_cfg5.json').read().strip().splitlines()[-1])
print(c.get('train_step_partitioned'))
PY
