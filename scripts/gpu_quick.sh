#!/bin/bash
python - <<'PY'
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
for dtype, s in ((torch.bfloat16, 2), (torch.float32, 4)):
    x = torch.randn(N, 256, device='cuda').to(dtype); out = torch.empty_like(x)
    bias = torch.zeros(256, device='cuda')
    for name, kw in (("gcn_fwd+bias", dict(row_scale=dinv, bias=bias)), ("gcn_fwd", dict(row_scale=dinv)), ("gcn_bwd", dict(row_scale=dinv, col_scale=dinv))):
        fn = lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, kw.get('row_scale'), kw.get('col_scale'), 0.0, None, kw.get('bias'), out=out, band=band)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/20
        alg = 2*N*256*s + 4*csr.nnz + 4*(N+1) + 4*N
        print(f"{dtype} {name}: {ms:.3f} ms {alg/ms/1e6/6553:.2%}", flush=True)
PY
PATHS=aggregate timeout 600 python scripts/gat_probe.py 2>&1 | grep "fwd "
PATHS=aggregate timeout 600 python scripts/tconv_probe.py 2>&1 | grep "fwd "
