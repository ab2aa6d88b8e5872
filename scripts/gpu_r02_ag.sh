#!/bin/bash
# round 2, call ag: segment softmax inside the fused attention kernel — tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_fused.py tests/test_gpu_layers.py tests/test_gpu_model.py tests/test_gpu_segsum.py -x -q -m gpu > gpurun_out/r02ag_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02ag_tests.log
for t in "" separate; do echo "B2G_GAT_SOFTMAX=$t"; B2G_GAT_SOFTMAX=$t timeout 200 python scripts/gatf_probe.py 2>&1 | grep -E "^fused|^unfused|diff" | head -3; done
