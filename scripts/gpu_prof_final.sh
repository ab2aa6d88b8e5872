#!/bin/bash
# Round evidence: launch list of the headline bench command, full captures of its two kernels, and of the GAT / BatchNorm kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"seg_rows|tc_linear" -s 4 -c 4 -f -o gpurun_out/prof_headline $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -n 2 gpurun_out/ncu_full.log
# GAT aggregate-first path (one fwd+bwd) and the FlowGNN train step (BatchNorm kernels)
PATHS=aggregate python scripts/gat_probe.py > gpurun_out/prof_gat_plain.log 2>&1 &&
PATHS=aggregate ncu --set full --clock-control none -k regex:"gatz_|rowdot8" -s 9 -c 4 -f -o gpurun_out/prof_gatz python scripts/gat_probe.py > gpurun_out/ncu_gatz.log 2>&1
echo "gatz capture exit $?"
FUSED=1 STEPS=1 python scripts/train_probe.py > gpurun_out/prof_train_plain.log 2>&1 &&
FUSED=1 STEPS=1 ncu --set full --clock-control none -k regex:"bn_" -s 30 -c 6 -f -o gpurun_out/prof_bn python scripts/train_probe.py > gpurun_out/ncu_bn.log 2>&1
echo "bn capture exit $?"
