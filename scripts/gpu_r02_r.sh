#!/bin/bash
# round 2, call r: ncu of the resident-plan 256 -> 1024 GEMM with the TMA-store epilogue
mkdir -p gpurun_out
ONLY=256x1024 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tc_linear_kernel" -s 3 -c 1 -o gpurun_out/r02r_gemm \
    python scripts/gemm_shapes_probe.py > gpurun_out/r02r_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r02r_ncu.log
