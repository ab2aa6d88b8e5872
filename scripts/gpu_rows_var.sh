#!/bin/bash
cat > /tmp/rows_var.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gnn_bfs_rans_b200 import ops, _lib
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
N = 250*200*200
o, n = hex_mesh_faces(250, 200, 200, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = Graph(ei, N)
csr = g.csr("sl", False); dinv = g.dinv(); band = g.band()
lib = _lib.load()
x = torch.randn(N, 256, device='cuda').bfloat16(); out = torch.empty_like(x)
for chunk in (32, 16):
    _lib.check(lib.b2g_set_seg_sched(chunk, 8192))
    fn = lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, dinv, None, 0.0, None, None, out=out, band=band)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"  chunk {chunk}: {e0.elapsed_time(e1)/20:.3f} ms", flush=True)
PY
for v in "" _w49 _w48 "" _w49; do echo "== variant ${v:-base}"; B2G_LIB=$PWD/gnn-bfs-rans_b200/libb2g$v.so python /tmp/rows_var.py; done
