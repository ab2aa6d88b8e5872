#!/bin/bash
mkdir -p gpurun_out
SEG_PROBE_ITERS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:seg_rows -s 1 -c 1 -f -o gpurun_out/seg_rows_full python scripts/seg_probe3.py > gpurun_out/seg_full.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/seg_full.log
