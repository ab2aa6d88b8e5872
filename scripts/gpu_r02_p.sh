#!/bin/bash
# round 2, call p: tensor-core tz_alpha — tests (attention files + layers + model), Transformer probe, ncu of the new kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_layers.py tests/test_gpu_model.py tests/test_gpu_segsum.py -x -q -m gpu > gpurun_out/r02p_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02p_tests.log
FWD_ONLY=1 PATHS=fused timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tz_alpha_mma|tc_linear_kernel" -c 6 -o gpurun_out/r02p_tz \
    python scripts/tconv_probe.py > gpurun_out/r02p_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r02p_ncu.log
