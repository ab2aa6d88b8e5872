#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gatz_bwd_src" -s 3 -c 1 -o gpurun_out/r02ac_bwdsrc \
    python scripts/bwdsrc_probe.py > gpurun_out/r02ac_ncu.log 2>&1; echo "ncu exit $?"
