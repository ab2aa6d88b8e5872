"""TransformerConv(256,256,heads=4,concat=False) bf16 on the cfg4 mesh: aggregate-first vs project-first."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_bfs_rans_b200 as b2g
from gnn_bfs_rans_b200 import ops
from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

nx, ny, nz = (int(v) for v in os.environ.get("MESH", "250,200,200").split(","))
N = nx * ny * nz
o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
ei = ops.build_graph_edges(o, n, 1, None, N, N)
F = 256


def timeit(fn, it=3, warm=1):
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


dtype = torch.bfloat16
torch.manual_seed(0)
layer = b2g.nn.TransformerConv(F, F, heads=4, concat=False).cuda().to(dtype).eval()
x = torch.randn(N, F, device='cuda').to(dtype)
gout = torch.randn(N, F, device='cuda').to(dtype)
outs = {}
for path in os.environ.get("PATHS", "aggregate,project").split(","):
    os.environ["B2G_TCONV_PATH"] = path
    with torch.no_grad():
        outs[path] = layer(x, ei)
        ms = timeit(lambda: layer(x, ei))
    if os.environ.get("FWD_ONLY"):
        print(f"{dtype} {path:9s}: fwd {ms:.2f} ms", flush=True)
        continue
    xg = x.clone().requires_grad_(True)

    def fb():
        xg.grad = None
        layer.zero_grad(set_to_none=True)
        layer(xg, ei).backward(gout)
    try:
        ms2 = timeit(fb, 2, 1)
    except Exception as e:
        ms2 = float('nan'); print("fwd+bwd failed:", str(e)[:300])
    print(f"{dtype} {path:9s}: fwd {ms:.2f} ms  fwd+bwd {ms2:.2f} ms  peak mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
    outs[path + "_gx"] = xg.grad.clone() if xg.grad is not None else None
    del xg
    torch.cuda.empty_cache()
if "project" in outs and "aggregate" in outs:
    d = (outs["aggregate"].float() - outs["project"].float()).abs().max() / outs["project"].float().abs().max()
    print("fwd rel diff aggregate vs project:", float(d))
    if outs["aggregate_gx"] is not None and outs["project_gx"] is not None:
        d = (outs["aggregate_gx"].float() - outs["project_gx"].float()).norm() / outs["project_gx"].float().norm()
        print("dx rel L2 diff:", float(d))
