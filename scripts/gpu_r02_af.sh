#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_layers.py -x -q -m gpu > gpurun_out/r02af_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02af_tests.log
for t in 1 0; do echo "B2G_GAT_ALPHA_PACKED=$t"; B2G_GAT_ALPHA_PACKED=$t timeout 200 python scripts/gatf_probe.py 2>&1 | grep -E "gat_alpha|^fused" | head -2; done
