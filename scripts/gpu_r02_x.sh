#!/bin/bash
# round 2, call x: the default bench line at N = 8 exactly as the driver launches it (weak scaling, halo_check, strong-scaling extra, cfg5 train step)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 \
  > gpurun_out/r02x_bench_n8.json 2> gpurun_out/r02x_bench_n8.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02x_bench_n8.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}); print(d.get('halo_check')); print(d.get('e2e'))
print(d.get('strong_scaling_cfg4')); print(d.get('train_step_partitioned')); print(d.get('roofline'))
PY
tail -3 gpurun_out/r02x_bench_n8.err
