#!/bin/bash
# 2-GPU: NCCL parity tests + weak/strong bench with the overlapped and the blocking halo exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_partition.py -q --tb=short 2>&1 | tail -15) > gpurun_out/r02e_partition_tests.log 2>&1
cat gpurun_out/r02e_partition_tests.log
for ov in 1 0; do
B2G_HALO_OVERLAP=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench2_ov$ov.json 2> gpurun_out/r02e_bench2_ov$ov.err
echo "bench ov=$ov rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r02e_bench2_ov$ov.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d.get(k) for k in ('value','ms_per_step','halo_check','strong_scaling_cfg4')}); print('e2e', d['e2e']['value'] if d.get('e2e') else None)
else:
    print(open('gpurun_out/r02e_bench2_ov$ov.err').read()[-1500:])
PY
done
