#!/bin/bash
# round 2, first GPU batch: GPU tests, the reference's own scripts through the drop-in, gradient-error diagnostic, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02a_pytest.log 2>&1
tail -3 gpurun_out/r02a_pytest.log
bash scripts/gpu_reference_dropin.sh > gpurun_out/r02a_refrun.log 2>&1
tail -12 gpurun_out/r02a_refrun.log
timeout 600 python scripts/grad_err_probe.py > gpurun_out/r02a_grad_err.log 2>&1
tail -5 gpurun_out/r02a_grad_err.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 1500 gpurun_out/r02a_bench.json
