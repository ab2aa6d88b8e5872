#!/bin/bash
# round 2, call s: attention backward on the tensor cores (gatz_bwd_dst / gatz_bwd_src) — tests, GAT / Transformer fwd+bwd A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_layers.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02s_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02s_tests.log
for mma in 1 0; do
  echo "B2G_ATTN_MMA=$mma"
  B2G_ATTN_MMA=$mma PATHS=fused timeout 300 python scripts/tconv_probe.py 2>&1 | tail -1
  B2G_ATTN_MMA=$mma timeout 300 python scripts/gatf_probe.py 2>&1 | grep -E "fused|unfused" | head -2
done > gpurun_out/r02s_ab.log 2>&1
cat gpurun_out/r02s_ab.log
