"""segw_gemm at cfg4, a few launches (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import graph_of
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
nx, ny, nz = (int(v) for v in os.environ.get("MESH", "250,200,200").split(","))
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
F = 256
torch.manual_seed(0)
x = torch.randn(N, F, device='cuda').bfloat16()
g = graph_of(ei, N)
csr = g.csr("sl", False)
dinv = g.dinv()
w = torch.randn(F, F, device='cuda').bfloat16() / 16
out = torch.empty(N, F, device='cuda', dtype=torch.bfloat16)
for _ in range(int(os.environ.get("REPS", "3"))):
    ops.segw_gemm(x, csr.rowptr, csr.col, N, w, None, col_scale=dinv, row_scale=dinv, band=g.band(), out=out)
torch.cuda.synchronize()
print("ok")
