#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -m gpu -q --tb=short --maxfail=6 -p no:cacheprovider 2>&1 | tail -8
timeout 600 python scripts/gat_probe.py 2>&1 | tail -12
