"""GCN backward aggregation at cfg4 (bf16 F = 256): per-entry-weighted CSR segment-sum over the transposed CSR,
B2G_ATTN_MMA=1 (seg_rows_wmma_kernel, mma.sync) vs 0 (seg_rows_kernel<kW>), and the unweighted forward kernel for reference."""
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
csr_t = g.csr("sl", True)
dinv = g.dinv()
x = torch.randn(N, 256, device='cuda').bfloat16()
out = torch.empty_like(x)


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


alg = 2 * N * 256 * 2 + 4 * csr_t.col.numel() + 4 * (N + 1) + 8 * N
ms = timeit(lambda: ops.seg_sum(x, csr_t.rowptr, csr_t.col, N, dinv, dinv, 0.0, None, None, out=out, band=g.band()))
print(f"weighted (GCN backward)  : {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s = {alg / ms / 1e6 / 6553:.3f} of HBM")
rows = torch.randint(0, N, (4096,), device='cuda')
rp = csr_t.rowptr.long()
ref = torch.zeros(4096, 256, dtype=torch.float64, device='cuda')
for k, i in enumerate(rows.tolist()[:512]):
    c = csr_t.col[rp[i]:rp[i + 1]].long()
    ref[k] = (dinv[i].double() * dinv[c].double().unsqueeze(1) * x[c].double()).sum(0)
err = (out[rows[:512]].double() - ref[:512]).abs().max() / ref[:512].abs().max()
print(f"sampled rel err vs fp64: {float(err):.2e}")
csr = g.csr("sl", False)
ms = timeit(lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, dinv, None, 0.0, None, None, out=out, band=g.band()))
print(f"unweighted (GCN forward) : {ms:.3f} ms")
