#!/bin/bash
# Round-2 ncu evidence (one gpurun call): launch list of the bench command; --set full of the headline kernels (seg_rows +
# tc_linear) with the DRAM traffic that bench.py reports as roofline.traffic (profiles/traffic.json is regenerated from THIS
# capture); --set full of the fused GAT forward kernels (rowdot8, gat_alpha, gatw_gemm) and of the unfused pair for comparison.
# The .ncu-rep files are summarised on the box (scripts/ncu_summary.py); only the fused kernel's report comes back.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/r02p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02p_ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none -k regex:"seg_rows|tc_linear" -s 4 -c 2 -f -o /tmp/prof_headline $CMD > gpurun_out/r02p_ncu_full.log 2>&1
echo "headline capture exit $?"
python scripts/ncu_summary.py /tmp/prof_headline.ncu-rep > gpurun_out/r02_sum_headline.txt 2>&1
REPS=3 python scripts/gatf_kernel_only.py > gpurun_out/r02p_gat_plain.log 2>&1 &&
REPS=3 ncu --set full --clock-control none --import-source on -k regex:"rowdot8|gat_alpha|gatw_gemm" -s 3 -c 3 -f -o gpurun_out/r02_gat_fused python scripts/gatf_kernel_only.py > gpurun_out/r02p_ncu_gat.log 2>&1
echo "gat fused capture exit $?"
python scripts/ncu_summary.py gpurun_out/r02_gat_fused.ncu-rep > gpurun_out/r02_sum_gat_fused.txt 2>&1
B2G_GAT_PATH=unfused PATHS=aggregate python scripts/gat_probe.py > gpurun_out/r02p_gatu_plain.log 2>&1 &&
B2G_GAT_PATH=unfused PATHS=aggregate ncu --set full --clock-control none -k regex:"gatz_fwd|tc_linear" -s 2 -c 2 -f -o /tmp/prof_gatu python scripts/gat_probe.py > gpurun_out/r02p_ncu_gatu.log 2>&1
echo "gat unfused capture exit $?"
python scripts/ncu_summary.py /tmp/prof_gatu.ncu-rep > gpurun_out/r02_sum_gat_unfused.txt 2>&1
grep -E "^---|time_duration|dram__bytes_(read|write).sum " gpurun_out/r02_sum_*.txt
ls -la gpurun_out/*.ncu-rep
