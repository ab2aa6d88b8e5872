"""gatz_bwd_src (4-head weighted row sums over the transposed CSR, bf16 F = 256) alone at cfg4: ms and GB/s; output strides
1024 (32-byte aligned rows) and 1032 (16-byte aligned rows, the GAT backward's [y | d a] operand)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import graph_of
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

nx, ny, nz = 250, 200, 200
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = graph_of(ei, N)
csr_t, perm = g.csr("sl", True), g.perm("sl")
nnz = csr_t.col.numel()
x = torch.randn(N, 256, device='cuda').bfloat16()
w = torch.rand(nnz, 4, device='cuda')


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


for ld in (1024, 1032):
    buf = torch.empty(N, ld, device='cuda', dtype=torch.bfloat16)
    out = buf[:, :1024]
    ms = timeit(lambda: ops.seg_wsum4(x, w, csr_t.rowptr, csr_t.col, perm, out, band=g.band()))
    alg = N * 256 * 2 + N * 1024 * 2 + 16 * nnz + 8 * nnz + 4 * N
    print(f"ld {ld}: {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s of algorithmic bytes", flush=True)
    del buf, out
