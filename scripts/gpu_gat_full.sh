#!/bin/bash
mkdir -p gpurun_out
PATHS=aggregate timeout 900 ncu --set full --import-source on --clock-control none -k regex:gatz_ -c 3 -f -o gpurun_out/gatz_full python scripts/gat_probe.py > gpurun_out/gatz_full.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/gatz_full.log
