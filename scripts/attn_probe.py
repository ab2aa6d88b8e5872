"""Kernel-only timing of the fused attention kernels (K4/K5) on the cfg4 mesh."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

nx, ny, nz = 250, 200, 200
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = Graph(ei, N)
H, F = 4, 256


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for dtype, s in ((torch.bfloat16, 2),):
    csr = g.csr("sl", False)
    xw = torch.empty(N, H * F, device='cuda', dtype=dtype).normal_()
    a = torch.randn(N, 2 * H, device='cuda')
    alg = N * H * F * s + N * F * s + 2 * 4 * N * H + 4 * csr.nnz + 4 * (N + 1)
    ms = timeit(lambda: ops.gat_fwd(xw, a, H, F, False, 0.2, csr.rowptr, csr.col, None, 0.0, 0, False, max_degree=g.max_degree("sl")))
    o1 = ops.gat_fwd(xw, a, H, F, False, 0.2, csr.rowptr, csr.col, None, 0.0, 0, False, max_degree=g.max_degree("sl"))[0]
    o0 = ops.gat_fwd(xw, a, H, F, False, 0.2, csr.rowptr, csr.col, None, 0.0, 0, False)[0]
    print(f"GAT fwd small-degree kernel: {ms:7.3f} ms ({alg/ms/1e6/6553:.1%}); max |diff| vs general kernel {float((o1.float()-o0.float()).abs().max()):.3e} (max |out| {float(o0.float().abs().max()):.2f})", flush=True)
    del o0, o1
    ms = timeit(lambda: ops.gat_fwd(xw, a, H, F, False, 0.2, csr.rowptr, csr.col, None, 0.0, 0, False))
    print(f"GAT fwd  {str(dtype):15s}: {ms:7.3f} ms  {alg/ms/1e6:6.0f} GB/s alg ({alg/ms/1e6/6553:.1%})  {csr.nnz/ms/1e6:.2f} G edges/s", flush=True)
    del xw, a
    csr = g.csr("raw", False)
    y = torch.empty(N, 3 * H * F + F, device='cuda', dtype=dtype).normal_()
    q, k, v, sk = y[:, :H * F], y[:, H * F:2 * H * F], y[:, 2 * H * F:3 * H * F], y[:, 3 * H * F:]
    ms = timeit(lambda: ops.tconv_fwd(q, k, v, sk, H, F, False, csr.rowptr, csr.col, 0.0, 0, False))
    alg = 3 * N * H * F * s + 2 * N * F * s + 4 * csr.nnz + 4 * (N + 1)
    print(f"Tconv fwd {str(dtype):15s}: {ms:7.3f} ms  {alg/ms/1e6:6.0f} GB/s alg ({alg/ms/1e6/6553:.1%})  {csr.nnz/ms/1e6:.2f} G edges/s", flush=True)
    del y, q, k, v, sk
    torch.cuda.empty_cache()
