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
x = torch.randn(N, 256, device='cuda').bfloat16()
def timeit(fn, it=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(it): fn()
    e1.record()
    t_host = (time.perf_counter() - t0) / it * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it, t_host
with torch.no_grad():
    print("eager   gpu ms %.3f  host enqueue ms %.3f" % timeit(lambda: layer(x, ei)))
    gf = b2g.graphs.GraphedForward(layer, x, ei)
    print("graphed gpu ms %.3f  host enqueue ms %.3f" % timeit(lambda: gf.graph.replay()))
