#!/bin/bash
# round 2, call u: plan 2 of the tcgen05 Linear (two row tiles per W k-block), identity rows_gather — tests, shapes A/B, layer probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_layers.py tests/test_gpu_gat_fused.py -x -q -m gpu > gpurun_out/r02u_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02u_tests.log
ONLY=1024x256,1032x256,3336x256 timeout 300 python scripts/gemm_shapes_probe.py > gpurun_out/r02u_gemm_dual.log 2>&1; echo "exit $?"; cat gpurun_out/r02u_gemm_dual.log
ONLY=1024x256,1032x256,3336x256 B2G_TC_DUAL=0 timeout 300 python scripts/gemm_shapes_probe.py > gpurun_out/r02u_gemm_single.log 2>&1; echo "exit $?"; cat gpurun_out/r02u_gemm_single.log
PATHS=fused timeout 300 python scripts/tconv_probe.py 2>&1 | tail -1
timeout 300 python scripts/gatf_probe.py 2>&1 | grep -E "fused|unfused" | head -2
