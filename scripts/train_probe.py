"""FlowGNN train step (fwd + loss + bwd + clip + Adam) on a hex mesh: time per step; used under ncu for the launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.flow_model import FlowGNN
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

nx, ny, nz = (int(v) for v in os.environ.get("MESH", "250,200,50").split(","))
lt = os.environ.get("LAYER", "GCN")
steps = int(os.environ.get("STEPS", "5"))
fused = os.environ.get("FUSED", "0") == "1"
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
torch.manual_seed(0)
kw = dict(fused_glue=True) if fused else {}
model = FlowGNN(3, 256, 7, 4, lt, dropout=0.1, **kw).cuda().to(torch.bfloat16).train()
opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5)
xin = torch.rand(N, 3, device='cuda', dtype=torch.bfloat16)
y = torch.rand(N, 7, device='cuda', dtype=torch.bfloat16)


def step():
    opt.zero_grad(set_to_none=True)
    loss = (model(xin, ei) - y).float().square().mean()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
print(f"{lt} fused={fused} N={N}: {e0.elapsed_time(e1)/steps:.2f} ms/step  loss {float(l):.4f}  peak {torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
