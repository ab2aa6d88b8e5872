"""gatw_gemm at cfg4, a few launches (for ncu)."""
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
torch.manual_seed(0)
layer = b2g.nn.GATConv(F, F, heads=4, concat=False).cuda().bfloat16().eval()
x = torch.randn(N, F, device='cuda').bfloat16()
g = graph_of(ei, N)
csr = g.csr("sl", False)
wc, v = layer._wc_v(torch.bfloat16)
wp = wc.view(F, H, F // 64, 64).permute(0, 2, 1, 3).reshape(F, H * F).contiguous()
with torch.no_grad():
    out = torch.empty(N, F, device='cuda', dtype=torch.bfloat16)
    for _ in range(int(os.environ.get("REPS", "4"))):          # the three kernels of the fused GATConv forward, in layer order
        if os.environ.get("SEPARATE"):                           # the three-kernel form (b2g_gat_alpha as its own launch)
            a = ops.rowdot8(x, v)
            alpha, _, _ = ops.gat_alpha(a, csr.rowptr, csr.col, H, 0.2, 0.0, 0, False)
            ops.gatw_gemm(x, csr.rowptr, csr.col, None, alpha, wp, layer.bias, N, H, band=g.band(), out=out)
        else:                                                    # the layer's default path: rowdot8 + gatw_gemm with the softmax inside
            out = layer(x, ei)
    torch.cuda.synchronize()
print("ok")
