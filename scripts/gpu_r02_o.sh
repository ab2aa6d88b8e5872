#!/bin/bash
# round 2, call o: PCIe link floor for the host pipeline + ncu of tz_alpha / the u GEMM (TransformerConv fused forward)
mkdir -p gpurun_out
timeout 300 python scripts/pcie_probe.py > gpurun_out/r02o_pcie.log 2>&1; echo "pcie exit $?"; cat gpurun_out/r02o_pcie.log
FWD_ONLY=1 PATHS=fused timeout 300 python scripts/tconv_probe.py > gpurun_out/r02o_tconv.log 2>&1; echo "tconv exit $?"; tail -4 gpurun_out/r02o_tconv.log
FWD_ONLY=1 PATHS=fused timeout 900 ncu --set full --import-source on --clock-control none -k regex:"tz_fwd_kernel|tc_linear_kernel<0>" -c 2 -o gpurun_out/r02o_tz \
    python scripts/tconv_probe.py > gpurun_out/r02o_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r02o_ncu.log
