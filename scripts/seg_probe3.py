"""seg_rows_kernel scheduling probe on the cfg4 mesh: chunk rows x panel rows x band hint; SEG_PROBE_ITERS=1 for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops, _lib
from gnn_bfs_rans_b200.graph import Graph
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

iters = int(os.environ.get("SEG_PROBE_ITERS", "10"))
quick = iters == 1
nx, ny, nz = 250, 200, 200
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
g = Graph(ei, N)
band = int((ei[0] - ei[1]).abs().max())
print("band", band, flush=True)
lib = _lib.load()
csr = g.csr("sl", False)
csr_raw = g.csr("raw", False)
dinv = g.dinv()
combos = [(64, 8192, 0), (32, 8192, band), (64, 8192, band), (128, 8192, band), (64, 4096, band), (64, 16384, band)]
if quick:
    combos = [(64, 8192, 0), (64, 8192, band)]
for dtype, s in ((torch.bfloat16, 2), (torch.float32, 4)):
    x = torch.randn(N, 256, device='cuda').to(dtype)
    out = torch.empty_like(x)
    for name, c, rs, cs, sc in (("gcn_fwd", csr, dinv, None, 0.0), ("gcn_bwd", csr, dinv, dinv, 0.0), ("gin", csr_raw, None, None, 1.0)):
        lib.b2g_set_seg_impl(1)
        ref = ops.seg_sum(x, c.rowptr, c.col, N, rs, cs, sc, None, None)
        lib.b2g_set_seg_impl(0)
        for chunk, panel, bnd in combos:
            _lib.check(lib.b2g_set_seg_sched(chunk, panel))
            fn = lambda: ops.seg_sum(x, c.rowptr, c.col, N, rs, cs, sc, None, None, out=out, band=bnd)
            for _ in range(2 if not quick else 0):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            err = (out.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
            ms = e0.elapsed_time(e1) / iters
            alg = 2 * N * 256 * s + 4 * c.nnz + 4 * (N + 1) + (4 * N if rs is not None else 0)
            print(f"{str(dtype):15s} {name:8s} chunk={chunk:4d} panel={panel:6d} band={bnd:6d}: {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s  ({alg/ms/1e6/6553:.2%})  relerr {err:.1e}", flush=True)
        if quick and name == "gcn_fwd":
            break
    del x, out, ref
