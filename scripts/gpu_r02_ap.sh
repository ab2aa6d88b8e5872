#!/bin/bash
# round 2, call ap: per-kernel table of the cfg5 share of one GPU (12.5 M cells, GAT L = 6, checkpointed train step)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r02ap_cfg5_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --train-step --train-checkpoint --train-cells 12500000 \
  > gpurun_out/r02ap_bench.json 2> gpurun_out/r02ap_bench.err; echo "ncu exit $?"
ls -la gpurun_out/r02ap_cfg5_launches.csv
