"""seg_sum micro-benchmark on the cfg4 mesh: achieved GB/s on algorithmic bytes for a few variants."""
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
from gnn_bfs_rans_b200 import _lib
for impl, dtype, s in ((1, torch.bfloat16, 2), (0, torch.bfloat16, 2), (1, torch.float32, 4), (0, torch.float32, 4)):
    _lib.load().b2g_set_seg_impl(impl)
    for F in (256, 128):
        x = torch.randn(N, F, device='cuda').to(dtype)
        out = torch.empty_like(x)
        for variant, use_dinv in (("sl", True), ("raw", False)):
            csr = g.csr(variant, False)
            dinv = g.dinv() if use_dinv else None
            fn = lambda: ops.seg_sum(x, csr.rowptr, csr.col, N, dinv, None, 0.0 if use_dinv else 1.0, None, None, out=out)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            if impl != 1 and use_dinv:   # same summation order -> bit-identical to the register-gather kernel
                ref = torch.empty_like(out)
                _lib.load().b2g_set_seg_impl(1)
                ops.seg_sum(x, csr.rowptr, csr.col, N, dinv, None, 0.0 if use_dinv else 1.0, None, None, out=ref)
                _lib.load().b2g_set_seg_impl(impl)
                torch.cuda.synchronize()
                assert torch.equal(ref, out), "bulk kernel differs from LDG kernel"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            alg = 2 * N * F * s + 4 * csr.nnz + 4 * (N + 1) + (4 * N if use_dinv else 0)
            print(f"impl{impl} {str(dtype):15s} F={F} {variant}: {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s algorithmic  ({alg/ms/1e6/6553:.2%} of measured HBM)  gather-model {csr.nnz*F*s/ms/1e6:.0f} GB/s", flush=True)
# pure streaming reference: copy N x 256 bf16
x = torch.randn(N, 256, device='cuda').bfloat16(); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"torch copy 5.12 GB: {ms:.3f} ms {2*x.numel()*2/ms/1e6:.0f} GB/s")
