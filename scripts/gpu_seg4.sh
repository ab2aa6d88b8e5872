#!/bin/bash
mkdir -p gpurun_out
export SEG_PROBE_ITERS=10
echo "== minb 4"; timeout 600 python scripts/seg_probe3.py 2>&1 | grep -v "chunk=  16\|panel= 32768\|chunk= 128\|panel= 16384" | tee gpurun_out/seg4_m4.log
echo "== minb 3"; B2G_LIB=$PWD/gnn-bfs-rans_b200/libb2g_m3.so timeout 600 python scripts/seg_probe3.py 2>&1 | grep -v "chunk=  16\|panel= 32768\|chunk= 128\|panel= 16384" | tee gpurun_out/seg4_m3.log
