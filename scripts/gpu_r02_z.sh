#!/bin/bash
# round 2, call z: CTA-pair variant of the fused attention kernel (tcgen05 cta_group::2) — kernel tests with a tight timeout, then timing
mkdir -p gpurun_out
B2G_GATW_PAIR=1 timeout 120 python -m pytest tests/test_gpu_gat_fused.py -x -q -m gpu -k "gatw or fused" > gpurun_out/r02z_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02z_tests.log
B2G_GATW_PAIR=1 timeout 120 python scripts/gatf_probe.py 2>&1 | grep -E "gatw_gemm band|fused|diff" | head -6
B2G_GATW_PAIR=0 timeout 120 python scripts/gatf_probe.py 2>&1 | grep -E "gatw_gemm band|fused" | head -3
