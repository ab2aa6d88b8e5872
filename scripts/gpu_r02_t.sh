#!/bin/bash
# round 2, call t: per-kernel times of the TransformerConv fused forward + backward (ncu launch list, cold-cache serialised)
mkdir -p gpurun_out
PATHS=fused timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r02t_tconv_launches.csv python scripts/tconv_probe.py > gpurun_out/r02t_ncu.log 2>&1; echo "ncu exit $?"
