#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -m gpu -q --tb=short --maxfail=6 -p no:cacheprovider -k "Transformer or dropout or flowgnn" 2>&1 | tail -40
MESH=${MESH:-250,200,100} timeout 600 python scripts/tconv_probe.py 2>&1 | tail -8
