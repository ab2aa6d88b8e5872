"""Runs one kernel configuration a few times (target of ncu captures).  usage: prof_kernel.py gemm|seg|gat|tconv [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

what = sys.argv[1]
from gnn_bfs_rans_b200 import _lib
_lib.load().b2g_set_seg_impl(int(os.environ.get("SEG_IMPL", "1")))
dtype = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "fp32") else torch.bfloat16
reps = 3
if what == "gemm":
    n, k, m = 2_000_000, 256, int(sys.argv[3]) if len(sys.argv) > 3 else 256
    x = torch.randn(n, k, device='cuda').to(dtype)
    w = (torch.randn(m, k, device='cuda') / 16).to(dtype)
    for _ in range(reps):
        ops.linear_fwd(x, w)
else:
    nx, ny, nz = 250, 200, 200
    N = nx * ny * nz
    o, nn = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei = ops.build_graph_edges(o, nn, 1, None, N, N)
    g = Graph(ei, N)
    F, H = 256, 4
    if what == "seg":
        csr = g.csr("sl", False)
        x = torch.randn(N, F, device='cuda').to(dtype)
        out = torch.empty_like(x)
        for _ in range(reps):
            ops.seg_sum(x, csr.rowptr, csr.col, N, g.dinv(), None, 0.0, None, None, out=out)
    elif what == "gat":
        csr = g.csr("sl", False)
        xw = torch.empty(N, H * F, device='cuda', dtype=dtype).normal_()
        a = torch.randn(N, 2 * H, device='cuda')
        for _ in range(reps):
            ops.gat_fwd(xw, a, H, F, False, 0.2, csr.rowptr, csr.col, None, 0.0, 0, False, max_degree=int(os.environ.get('MAXDEG', '7')))
    elif what == "tconv":
        csr = g.csr("raw", False)
        y = torch.randn(N, 3 * H * F + F, device='cuda').to(dtype)
        q, k, v, sk = y[:, :H * F], y[:, H * F:2 * H * F], y[:, 2 * H * F:3 * H * F], y[:, 3 * H * F:]
        for _ in range(reps):
            ops.tconv_fwd(q, k, v, sk, H, F, False, csr.rowptr, csr.col, 0.0, 0, False)
torch.cuda.synchronize()
print("done", what)
