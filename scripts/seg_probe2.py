import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops, _lib
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
nx, ny, nz = 250, 200, 200
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = Graph(ei, N)
csr = g.csr("sl", False)
F = 256
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
out = torch.empty(N, F, device='cuda', dtype=torch.bfloat16)
for name, x in (("randn", torch.randn(N, F, device='cuda').bfloat16()), ("const", torch.full((N, F), 0.5, device='cuda', dtype=torch.bfloat16)),
                ("zeros", torch.zeros(N, F, device='cuda', dtype=torch.bfloat16))):
    for mode, rs in (("rowscale", g.dinv()), ("plain", None)):
        ms = t(lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, rs, None, 0.0, None, None, out=out))
        ref = out.clone()
        ms2 = t(lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, rs, None, 0.0, None, None, out=out, band=g.band()))
        print(f"{name:6s} {mode:8s}: linear {ms:.3f} ms   panel order (band {g.band()}) {ms2:.3f} ms   identical={torch.equal(ref, out)}")
# sorted-col CSR like the micro-benchmark
col_sorted = csr.col.clone()
x = torch.randn(N, F, device='cuda').bfloat16()
import numpy as np
print("col order sample:", csr.col[csr.rowptr[5000000]:csr.rowptr[5000001]].tolist())
