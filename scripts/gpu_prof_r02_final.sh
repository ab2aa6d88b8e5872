#!/bin/bash
# End of round 2: the launch list of the bench command and --set full of the headline kernels and of the fused GAT forward with the
# FINAL code (the tcgen05 Linear and the fused attention kernel changed after scripts/gpu_prof_r02.sh ran).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
timeout 300 $CMD > gpurun_out/r02f_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu_launch.log 2>&1
echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none -k regex:"seg_rows|tc_linear" -s 4 -c 2 -f -o /tmp/prof_headline $CMD > gpurun_out/r02f_ncu_full.log 2>&1
echo "headline capture exit $?"
python scripts/ncu_summary.py /tmp/prof_headline.ncu-rep > gpurun_out/r02f_sum_headline.txt 2>&1
REPS=3 timeout 200 python scripts/gatf_kernel_only.py > gpurun_out/r02f_gat_plain.log 2>&1 &&
REPS=3 timeout 400 ncu --set full --clock-control none -k regex:"rowdot8|gat_alpha|gatw_gemm" -s 2 -c 2 -f -o /tmp/prof_gat python scripts/gatf_kernel_only.py > gpurun_out/r02f_ncu_gat.log 2>&1
echo "gat fused capture exit $?"
python scripts/ncu_summary.py /tmp/prof_gat.ncu-rep > gpurun_out/r02f_sum_gat_fused.txt 2>&1
grep -E "^---|time_duration|dram__bytes_(read|write).sum " gpurun_out/r02f_sum_*.txt
