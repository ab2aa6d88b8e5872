#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/r02k_smoke.log
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) | tee gpurun_out/r02k_pytest.log
