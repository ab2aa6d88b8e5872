"""Fused GATConv forward (csrc/gat_fused.cu) at cfg4: per-kernel CUDA-event times (rowdot8, gat_alpha, gatw_gemm) and the
layer forward / forward+backward, fused vs unfused (B2G_GAT_PATH)."""
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
F, H = 256, 4


def timeit(fn, it=5, warm=2):
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
layer = b2g.nn.GATConv(F, F, heads=4, concat=False).cuda().bfloat16().eval()
x = torch.randn(N, F, device='cuda').bfloat16()
g = graph_of(ei, N)
csr = g.csr("sl", False)
wc, v = layer._wc_v(torch.bfloat16)
wp = wc.view(F, H, F // 64, 64).permute(0, 2, 1, 3).reshape(F, H * F).contiguous()
with torch.no_grad():
    a = ops.rowdot8(x, v)
    alpha, _, _ = ops.gat_alpha(a, csr.rowptr, csr.col, H, 0.2, 0.0, 0, False)
    out = torch.empty(N, F, device='cuda', dtype=torch.bfloat16)
    print(f"rowdot8   {timeit(lambda: ops.rowdot8(x, v)):.3f} ms")
    print(f"gat_alpha {timeit(lambda: ops.gat_alpha(a, csr.rowptr, csr.col, H, 0.2, 0.0, 0, False)):.3f} ms")
    for band in (g.band(), 0):
        ms = timeit(lambda: ops.gatw_gemm(x, csr.rowptr, csr.col, None, alpha, wp, layer.bias, N, H, band=band, out=out))
        alg = 2 * N * F * 2 + 4 * csr.nnz + 16 * csr.nnz + 4 * (N + 1)
        print(f"gatw_gemm band={band}: {ms:.3f} ms  algorithmic {alg / 1e9:.2f} GB -> {alg / ms / 1e6:.0f} GB/s", flush=True)
gout = torch.randn(N, F, device='cuda').bfloat16()
outs = {}
for path in ("", "unfused"):
    os.environ["B2G_GAT_PATH"] = path
    with torch.no_grad():
        outs[path] = layer(x, ei)
        ms = timeit(lambda: layer(x, ei))
    xg = x.clone().requires_grad_(True)

    def fb():
        xg.grad = None
        layer.zero_grad(set_to_none=True)
        layer(xg, ei).backward(gout)
    torch.cuda.reset_peak_memory_stats()
    ms2 = timeit(fb, 3, 1)
    print(f"{path or 'fused':8s}: fwd {ms:.2f} ms  fwd+bwd {ms2:.2f} ms  peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
    del xg
    torch.cuda.empty_cache()
d = (outs[""].float() - outs["unfused"].float()).abs().max() / outs["unfused"].float().abs().max()
print("fwd rel diff fused vs unfused:", float(d))
