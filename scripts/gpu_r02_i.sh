#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 250 python -m pytest tests/test_gpu_gat_fused.py -q --tb=short -x 2>&1 | tail -8) > gpurun_out/r02i_tests.log 2>&1
cat gpurun_out/r02i_tests.log
PATHS=fused,unfused timeout 200 python scripts/tconv_probe.py 2>&1 | grep -E "fwd|rror" | tee gpurun_out/r02i_tconv_probe.log
timeout 100 python scripts/gatf_probe.py 2>&1 | tail -5
