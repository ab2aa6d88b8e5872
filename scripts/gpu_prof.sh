#!/bin/bash
# ncu evidence for the headline command (B200_PROFILING.md recipe): plain run first, then launch list + full capture.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-seg_sum} -s 3 -c 2 -f -o gpurun_out/prof_${TAG:-segsum} $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -n 5 gpurun_out/ncu_full.log
