#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_kernel.py gat > gpurun_out/p_gat.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gat_fwd_small -s 1 -c 1 -f -o gpurun_out/prof_gat python scripts/prof_kernel.py gat > gpurun_out/ncu_gat.log 2>&1
echo "gat exit $?"; tail -n 3 gpurun_out/ncu_gat.log
