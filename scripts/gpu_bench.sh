#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --small --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small exit $?"; tail -c 1500 gpurun_out/bench_small.json; tail -n 5 gpurun_out/bench_small.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full exit $?"; cat gpurun_out/bench_full.json; tail -n 5 gpurun_out/bench_full.err
