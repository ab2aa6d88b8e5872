#!/bin/bash
# round 2, call aj (2 GPUs): partition tests incl. the NCCL ones, then the default N = 2 bench line as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r02aj_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02aj_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02aj_bench_n2.json 2> gpurun_out/r02aj_bench_n2.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02aj_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}); print(d.get('halo_check', {}).get('ok')); print(d['e2e']['ms_per_step'])
print(d.get('strong_scaling_cfg4'))
PY
