#!/bin/bash
# round 2, call v: the whole GPU suite, smoke(), the default bench line and the reference arm on one B200
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02v_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02v_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02v_smoke.log
timeout 900 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d.get('parity_check'))
ex = d.get('extras', {})
for k, v in ex.items():
    print(k, json.dumps(v)[:400])
r = json.loads(open('gpurun_out/r02v_bench_ref.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline'))
PY
