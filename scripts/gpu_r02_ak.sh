#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_layers.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02ak_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02ak_tests.log
PATHS=fused timeout 300 python scripts/tconv_probe.py 2>&1 | tail -1
