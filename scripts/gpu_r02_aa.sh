#!/bin/bash
# round 2, call aa: GCN backward aggregation on mma.sync — tests, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_segsum.py tests/test_gpu_layers.py tests/test_gpu_model.py tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r02aa_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02aa_tests.log
for mma in 1 0; do echo "B2G_ATTN_MMA=$mma"; B2G_ATTN_MMA=$mma timeout 200 python scripts/segw_probe.py 2>&1 | tail -3; done | tee gpurun_out/r02aa_segw.log
