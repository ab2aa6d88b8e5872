#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_training.py tests/test_gpu_model.py -q --tb=short 2>&1 | tail -30) > gpurun_out/r02f_training_tests.log 2>&1
cat gpurun_out/r02f_training_tests.log
