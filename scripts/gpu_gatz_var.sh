#!/bin/bash
for v in "" _v42 _v43 _v83; do
  echo "== variant ${v:-base}"
  B2G_LIB=$PWD/gnn-bfs-rans_b200/libb2g$v.so PATHS=aggregate timeout 600 python scripts/gat_probe.py 2>&1 | grep "fwd "
done
