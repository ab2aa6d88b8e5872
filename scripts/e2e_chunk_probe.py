import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_bfs_rans_b200 as b2g
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
N = 250 * 200 * 200
o, n = hex_mesh_faces(250, 200, 200, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
layer = b2g.nn.GCNConv(256, 256).cuda().to(torch.bfloat16).eval()
hx = torch.randn(N, 256).bfloat16().pin_memory()
hei = ei.cpu().pin_memory()
hout = torch.empty(N, 256, dtype=torch.bfloat16).pin_memory()
for R in (1 << 19, 1 << 18, 1 << 17, 1 << 20):
    for _ in range(2):
        b2g.streaming.gcn_forward_host(layer, hx, hei, hout, rows_per_chunk=R)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        b2g.streaming.gcn_forward_host(layer, hx, hei, hout, rows_per_chunk=R)
    torch.cuda.synchronize()
    print(f"rows_per_chunk {R:8d}: {(time.perf_counter() - t0) / 4 * 1e3:.1f} ms", flush=True)
