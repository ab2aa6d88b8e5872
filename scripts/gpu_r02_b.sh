#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 150 python -m pytest tests/test_gpu_gat_fused.py -x -q 2>&1 | tail -25) > gpurun_out/r02b_fused_tests.log 2>&1
cat gpurun_out/r02b_fused_tests.log
timeout 120 python scripts/gatf_probe.py > gpurun_out/r02b_gatf_probe.log 2>&1
cat gpurun_out/r02b_gatf_probe.log
(timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py tests/test_gpu_partition.py tests/test_gpu_segsum.py -q 2>&1 | tail -30) > gpurun_out/r02b_layer_tests.log 2>&1
cat gpurun_out/r02b_layer_tests.log
