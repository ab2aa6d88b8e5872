#!/bin/bash
mkdir -p gpurun_out
for pr in 0 1; do echo "B2G_GATW_PAIR=$pr"; B2G_GATW_PAIR=$pr timeout 200 python scripts/gatf_probe.py 2>&1 | grep -E "gatw_gemm band|^fused" | head -3; done
