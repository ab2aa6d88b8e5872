#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_segsum.py tests/test_gpu_layers.py tests/test_gpu_model.py -m gpu -q --tb=short --maxfail=6 -p no:cacheprovider 2>&1 | tail -5
bash scripts/gpu_quick.sh
