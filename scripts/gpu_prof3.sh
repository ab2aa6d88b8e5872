#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_kernel.py seg > gpurun_out/p_seg.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:seg_sum -s 1 -c 1 -f -o gpurun_out/prof_seg4 python scripts/prof_kernel.py seg > gpurun_out/ncu_seg.log 2>&1
echo "seg exit $?"
python scripts/prof_kernel.py gemm > gpurun_out/p_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_linear -s 1 -c 1 -f -o gpurun_out/prof_gemm3 python scripts/prof_kernel.py gemm > gpurun_out/ncu_gemm.log 2>&1
echo "gemm exit $?"
