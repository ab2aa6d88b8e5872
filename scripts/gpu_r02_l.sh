#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_gcn_fused.py -q --tb=line 2>&1 | tail -5) | tee gpurun_out/r02l_tests.log
timeout 100 python scripts/gcnf_probe.py 2>&1 | tail -12 | tee gpurun_out/r02l_gcnf_probe.log
