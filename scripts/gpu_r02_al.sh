#!/bin/bash
# round 2, call al: ReLU backward fused into the dgrad GEMM (GIN MLP as one autograd node) — tests, GIN fwd+bwd A/B, GEMM check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_layers.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02al_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02al_tests.log
for m in "" split "" split; do echo "B2G_GIN_MLP=$m"; B2G_GIN_MLP=$m timeout 200 python scripts/gin_probe.py 2>&1 | tail -1; done
ONLY=256x256,256x1024 timeout 200 python scripts/gemm_shapes_probe.py 2>&1 | tail -2
