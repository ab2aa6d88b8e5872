#!/bin/bash
# Runs each GPU test file in its own process (a CUDA fault in one must not poison the rest) and keeps logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $n exit $?"; tail -n 25 gpurun_out/$n.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 12 gpurun_out/smoke.log
