#!/bin/bash
# round 2, call y (2 GPUs): rowdot8 on mma.sync, SM reserve for the overlapped halo exchange — tests, then N = 2 A/B of the reserve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_partition.py tests/test_gpu_linear.py -x -q -m gpu > gpurun_out/r02y_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02y_tests.log
timeout 200 python scripts/gatf_probe.py 2>&1 | grep -E "rowdot8|gat_alpha|gatw_gemm band|fused" | head -6
for r in 0 8 16; do
  B2G_HALO_RESERVE_SMS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((r % 10)) \
     bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/r02y_n2_reserve$r.json 2> gpurun_out/r02y_n2_reserve$r.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02y_n2_reserve$r.json').read().strip().splitlines()[-1]); print('reserve $r', d['ms_per_step'], d['value'], d['halo_check']['ok'], d['roofline']['kernel_ms'])"
done
