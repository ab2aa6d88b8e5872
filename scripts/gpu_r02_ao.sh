#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gatz_bwd_dst_mma" -s 1 -c 1 -o gpurun_out/r02ao_bwddst \
    python scripts/gatf_probe.py > gpurun_out/r02ao_ncu.log 2>&1; echo "ncu exit $?"
