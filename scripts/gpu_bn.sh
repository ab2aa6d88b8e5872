#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_bn.py tests/test_gpu_model.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider 2>&1 | tail -40
timeout 600 python scripts/train_probe.py 2>&1 | tail -1
FUSED=1 timeout 600 python scripts/train_probe.py 2>&1 | tail -1
LAYER=GAT FUSED=1 timeout 600 python scripts/train_probe.py 2>&1 | tail -1
