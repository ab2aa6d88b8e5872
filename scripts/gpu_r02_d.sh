#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scripts/gatf_kernel_only.py > gpurun_out/r02d_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gatw_gemm -s 2 -c 1 -f -o gpurun_out/r02d_gatw python scripts/gatf_kernel_only.py > gpurun_out/r02d_ncu.log 2>&1
echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/r02d_gatw.ncu-rep > gpurun_out/r02d_sum.txt 2>&1
cat gpurun_out/r02d_sum.txt
ls -la gpurun_out/r02d_gatw.ncu-rep
