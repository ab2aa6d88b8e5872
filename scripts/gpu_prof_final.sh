#!/bin/bash
# Round evidence: launch list of the headline bench command + full captures of its two kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"seg_sum_rows|tc_linear" -s 4 -c 4 -f -o gpurun_out/prof_headline $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -n 3 gpurun_out/ncu_full.log
