#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gat_fused.py -x -q -m gpu -k "weighted_row_sums or gatconv or transformerconv" > gpurun_out/r02ab_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02ab_tests.log
timeout 200 python scripts/bwdsrc_probe.py 2>&1 | tail -2
B2G_ATTN_MMA=0 timeout 200 python scripts/bwdsrc_probe.py 2>&1 | tail -2
