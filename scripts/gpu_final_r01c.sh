#!/bin/bash
# End-of-round check + refreshed ncu evidence for the headline kernels (after the seg_rows epilogue change):
# GPU test suite, launch list and one --set full capture of the bench command, the full bench line, smoke().
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"seg_rows|tc_linear" -s 4 -c 2 -f -o /tmp/prof_headline $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
python scripts/ncu_summary.py /tmp/prof_headline.ncu-rep > gpurun_out/r01c_sum_headline.txt 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 200 gpurun_out/bench.err
python __graft_entry__.py smoke 2>&1 | tail -2
