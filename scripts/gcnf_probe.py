"""Fused GCNConv / GINConv forward (csrc/gcn_fused.cu) at cfg4: kernel and layer timings, fused vs unfused."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_bfs_rans_b200 as b2g
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import graph_of
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
nx, ny, nz = (int(v) for v in os.environ.get("MESH", "250,200,200").split(","))
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
F = 256


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


torch.manual_seed(0)
x = torch.randn(N, F, device='cuda').bfloat16()
g = graph_of(ei, N)
for kind in ("GCN", "GIN"):
    layer = (b2g.nn.GCNConv(F, F) if kind == "GCN" else
             b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F)))).cuda().bfloat16().eval()
    outs = {}
    for path in ("fused", "unfused"):
        os.environ["B2G_GCN_PATH"] = path
        os.environ["B2G_GIN_PATH"] = path
        with torch.no_grad():
            outs[path] = layer(x, ei)
            ms = timeit(lambda: layer(x, ei))
        print(f"{kind} {path:8s}: fwd {ms:.3f} ms", flush=True)
    d = (outs["fused"].float() - outs["unfused"].float()).abs().max() / outs["unfused"].float().abs().max()
    print(f"{kind} fwd rel diff fused vs unfused: {float(d):.2e}")
csr = g.csr("sl", False)
dinv = g.dinv()
w = torch.randn(F, F, device='cuda').bfloat16() / 16
out = torch.empty(N, F, device='cuda', dtype=torch.bfloat16)
for band in (g.band(), 0):
    ms = timeit(lambda: ops.segw_gemm(x, csr.rowptr, csr.col, N, w, None, col_scale=dinv, row_scale=dinv, band=band, out=out))
    alg = 2 * N * F * 2 + 4 * csr.nnz + 4 * (N + 1) + 4 * N
    print(f"segw_gemm band={band}: {ms:.3f} ms  algorithmic {alg / 1e9:.2f} GB -> {alg / ms / 1e6:.0f} GB/s = {alg / ms / 1e6 / 6553:.2%} of HBM", flush=True)
