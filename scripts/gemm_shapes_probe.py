"""tcgen05 Linear (gemm_tc.cu) at cfg4 row counts for the shapes the layers use: ms, GB/s of algorithmic bytes.
B2G_TC_TMA_STORE=0 selects the LDS + row-store epilogue for A/B runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops

N = int(os.environ.get("ROWS", 10_000_000))
SHAPES = [(256, 256, True, True), (256, 1024, True, False), (1024, 256, True, False), (256, 128, True, False), (256, 1280, False, False), (1032, 256, False, False), (3336, 256, False, False)]
if os.environ.get("ONLY"):
    SHAPES = [s_ for s_ in SHAPES if f"{s_[0]}x{s_[1]}" in os.environ["ONLY"].split(",")]
for (k, m, bias, rs) in SHAPES:
    x = torch.empty(N, k, device='cuda', dtype=torch.bfloat16)
    for r0 in range(0, N, 1 << 20):
        x[r0:r0 + (1 << 20)] = torch.randn(min(1 << 20, N - r0), k, device='cuda').bfloat16()
    w = (torch.randn(m, k, device='cuda') / k ** 0.5).bfloat16()
    b = torch.randn(m, device='cuda') if bias else None
    r = torch.rand(N, device='cuda') if rs else None
    y = torch.empty(N, m, device='cuda', dtype=torch.bfloat16)
    for _ in range(2):
        ops.linear_fwd(x, w, b, row_scale=r, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.linear_fwd(x, w, b, row_scale=r, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rows = torch.randint(0, N, (2048,), device='cuda')
    ref = x[rows].double() @ w.double().T
    if rs:
        ref = ref * r[rows].double().unsqueeze(1)
    if bias:
        ref = ref + b.double()
    err = float((y[rows].double() - ref).abs().max() / ref.abs().max())
    print(f"{k:5d} -> {m:5d}: {ms:7.3f} ms  {(N * (k + m) * 2) / ms / 1e6:6.0f} GB/s  sampled rel err {err:.2e}", flush=True)
    del x, y
