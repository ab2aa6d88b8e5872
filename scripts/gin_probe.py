"""GINConv(256->256->256) bf16 fwd+bwd on the cfg4 mesh (for the ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_bfs_rans_b200 as b2g
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
N = 250 * 200 * 200
o, n = hex_mesh_faces(250, 200, 200, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
F = 256
kind = os.environ.get("KIND", "GIN")
torch.manual_seed(0)
if kind == "GIN":
    layer = b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F)))
else:
    layer = b2g.nn.GCNConv(F, F)
layer = layer.cuda().to(torch.bfloat16).eval()
x = torch.randn(N, F, device='cuda').bfloat16().requires_grad_(True)
g = torch.randn(N, F, device='cuda').bfloat16()
for _ in range(3):
    x.grad = None
    layer.zero_grad(set_to_none=True)
    layer(x, ei).backward(g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    x.grad = None
    layer.zero_grad(set_to_none=True)
    layer(x, ei).backward(g)
e1.record(); torch.cuda.synchronize()
print(kind, "fwd+bwd ms", e0.elapsed_time(e1) / 3)
