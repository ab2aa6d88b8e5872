#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_gat_fused.py -x -q 2>&1 | tail -25) > gpurun_out/r02b_fused_tests.log 2>&1
cat gpurun_out/r02b_fused_tests.log
timeout 100 python scripts/gatf_probe.py > gpurun_out/r02b_gatf_probe.log 2>&1
cat gpurun_out/r02b_gatf_probe.log
(timeout 200 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -q --tb=line 2>&1 | grep -E "Error|assert|FAILED|passed|failed" | cut -c1-300) > gpurun_out/r02b_layer_tests.log 2>&1
cat gpurun_out/r02b_layer_tests.log
