#!/bin/bash
# Round evidence: launch list of the headline bench command, full captures of its two kernels and of the attention kernels.
# The .ncu-rep files are summarised on the box (scripts/ncu_summary.py) and removed: gpurun returns at most 64 MiB.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none -k regex:"seg_rows|tc_linear" -s 4 -c 2 -f -o /tmp/prof_headline $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
python scripts/ncu_summary.py /tmp/prof_headline.ncu-rep > gpurun_out/sum_headline.txt 2>&1
PATHS=aggregate python scripts/gat_probe.py > gpurun_out/prof_gat_plain.log 2>&1 &&
PATHS=aggregate ncu --set full --clock-control none -k regex:"gatz_" -s 9 -c 3 -f -o /tmp/prof_gatz python scripts/gat_probe.py > gpurun_out/ncu_gatz.log 2>&1
echo "gatz capture exit $?"
python scripts/ncu_summary.py /tmp/prof_gatz.ncu-rep > gpurun_out/sum_gatz.txt 2>&1
PATHS=aggregate ncu --set full --clock-control none -k regex:"tz_fwd|gatz_bwd_dst" -s 2 -c 2 -f -o /tmp/prof_tz python scripts/tconv_probe.py > gpurun_out/ncu_tz.log 2>&1
echo "tz capture exit $?"
python scripts/ncu_summary.py /tmp/prof_tz.ncu-rep > gpurun_out/sum_tz.txt 2>&1
ls -la gpurun_out | head -20
