#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scripts/gcnf_kernel_only.py > gpurun_out/r02m_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segw_gemm -s 1 -c 1 -f -o gpurun_out/r02m_segw python scripts/gcnf_kernel_only.py > gpurun_out/r02m_ncu.log 2>&1
echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/r02m_segw.ncu-rep > gpurun_out/r02m_sum.txt 2>&1
grep -E "time_duration|dram__bytes_(read|write).sum |issue_active|l1tex__throughput|lts__throughput|tensor_cycles|hit_rate" gpurun_out/r02m_sum.txt
