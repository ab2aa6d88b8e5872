#!/bin/bash
# round 2, call q: TMA-store epilogue of the tcgen05 Linear — tests, shapes A/B, layer probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_layers.py tests/test_gpu_gat_fused.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02q_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02q_tests.log
timeout 300 python scripts/gemm_shapes_probe.py > gpurun_out/r02q_gemm_tma.log 2>&1; echo "exit $?"; cat gpurun_out/r02q_gemm_tma.log
B2G_TC_TMA_STORE=0 timeout 300 python scripts/gemm_shapes_probe.py > gpurun_out/r02q_gemm_rowstore.log 2>&1; echo "exit $?"; cat gpurun_out/r02q_gemm_rowstore.log
PATHS=fused timeout 300 python scripts/tconv_probe.py > gpurun_out/r02q_tconv.log 2>&1; echo "tconv exit $?"; tail -2 gpurun_out/r02q_tconv.log
