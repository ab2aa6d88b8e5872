"""GCN layer forward: one GEMM + one aggregation (baseline) vs row-chunked interleaving so that the aggregation reads the
projected rows from L2 (experiment)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_bfs_rans_b200 as b2g
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
N = 250 * 200 * 200
o, n = hex_mesh_faces(250, 200, 200, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = Graph(ei, N)
csr = g.csr("sl", False); dinv = g.dinv(); band = g.band_raw()
F = 256
torch.manual_seed(0)
w = (torch.randn(F, F, device='cuda') * 0.05).bfloat16()
bias = torch.zeros(F, device='cuda')
x = torch.randn(N, F, device='cuda').bfloat16()
xs = torch.empty_like(x); out = torch.empty_like(x)

def base():
    ops.linear_fwd(x, w, None, row_scale=dinv, out=xs)
    ops.seg_sum(xs, csr.rowptr, csr.col, N, dinv, None, 0.0, None, bias, out=out, band=g.band())

def chunked(R):
    bounds = [(r0, min(r0 + R, N)) for r0 in range(0, N, R)]
    nxt = 0
    for (r0, r1) in bounds:
        ops.linear_fwd(x[r0:r1], w, None, row_scale=dinv[r0:r1], out=xs[r0:r1])
        while nxt < len(bounds) and min(bounds[nxt][1] + band, N) <= r1:
            a0, a1 = bounds[nxt]
            ops.seg_sum(xs, csr.rowptr[a0:a1 + 1], csr.col, a1 - a0, dinv[a0:a1], None, 0.0, None, bias, out=out[a0:a1], band=g.band())
            nxt += 1
    while nxt < len(bounds):
        a0, a1 = bounds[nxt]
        ops.seg_sum(xs, csr.rowptr[a0:a1 + 1], csr.col, a1 - a0, dinv[a0:a1], None, 0.0, None, bias, out=out[a0:a1], band=g.band())
        nxt += 1

def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

base(); ref = out.clone()
print("baseline      %.3f ms" % timeit(base))
for R in (1 << 20, 1 << 18, 1 << 17, 1 << 16, 1 << 15):
    chunked(R); ok = torch.equal(out, ref)
    ms = timeit(lambda: chunked(R))
    # same launches replayed from a CUDA graph (no launch gaps)
    gph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chunked(R)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(gph):
        chunked(R)
    msg = timeit(lambda: gph.replay())
    print("chunk %8d  eager %.3f ms  graph %.3f ms  identical=%s" % (R, ms, msg, ok), flush=True)
