#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_segsum.py -m gpu -q -x --tb=short -k "host_pipelined" -p no:cacheprovider 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
