#!/bin/bash
# round 2, call ad: TMA-store epilogue of the fused attention kernel — tests, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_segsum.py -x -q -m gpu > gpurun_out/r02ad_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02ad_tests.log
for t in 1 0; do echo "B2G_GATW_TMA_STORE=$t"; B2G_GATW_TMA_STORE=$t timeout 200 python scripts/gatf_probe.py 2>&1 | grep -E "gatw_gemm band=5|^fused" | head -2; B2G_GATW_TMA_STORE=$t FWD_ONLY=1 PATHS=fused timeout 200 python scripts/tconv_probe.py 2>&1 | tail -1; done
