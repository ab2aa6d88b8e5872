#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gat_fused.py -x -q -m gpu -k "softmax_inside" > gpurun_out/r02ak_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02ak_tests.log
