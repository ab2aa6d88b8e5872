#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -q --tb=line 2>&1 | grep -E "Error|assert|FAILED|passed|failed" | cut -c1-400) > gpurun_out/r02c_fail.log 2>&1
cat gpurun_out/r02c_fail.log
