#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_segsum.py tests/test_gpu_gat_fused.py tests/test_gpu_layers.py -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r02g_tests.log 2>&1
cat gpurun_out/r02g_tests.log
