"""First-contact probe of the tcgen05 GEMM: small shapes, compare with torch, print timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_bfs_rans_b200 import ops, _lib

lib = _lib.load()
print("impl for (1000,256,256,bf16):", lib.b2g_linear_impl(1000, 256, 256, 1, 0), flush=True)
torch.manual_seed(0)
for (n, k, m) in [(128, 64, 256), (128, 256, 256), (1000, 256, 256), (4099, 256, 1032), (300, 72, 40), (100000, 256, 256)]:
    x = torch.randn(n, k, device='cuda').bfloat16()
    w = (torch.randn(m, k, device='cuda') / k ** 0.5).bfloat16()
    b = torch.randn(m, device='cuda')
    ops.GEMM_IMPL = 2
    y, _ = ops.linear_fwd(x, w, b, act=0)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T + b.double()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"n={n} k={k} m={m}: rel err {err:.3e}", flush=True)
    if m >= 64:
        mm = (m // 8) * 8 - 8
        y2, aux = ops.linear_fwd(x, w, None, m_main=mm)
        ref2 = x.double() @ w.double().T
        e1 = float((y2.double() - ref2[:, :mm]).abs().max() / ref2.abs().max())
        e2 = float((aux.double() - ref2[:, mm:]).abs().max() / ref2.abs().max())
        print(f"   aux split at {mm}: main {e1:.3e} aux {e2:.3e}", flush=True)
n, k, m = 2_000_000, 256, 256
x = torch.randn(n, k, device='cuda').bfloat16()
w = (torch.randn(m, k, device='cuda') / 16).bfloat16()
for impl in (2, 1):
    ops.GEMM_IMPL = impl
    for _ in range(2):
        ops.linear_fwd(x, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.linear_fwd(x, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"impl {impl}: {ms:.3f} ms  {2*n*k*m/ms/1e9:.1f} TFLOP/s  {(n*k*2+n*m*2)/ms/1e6:.0f} GB/s", flush=True)
a = torch.matmul(x, w.t()); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    torch.matmul(x, w.t())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"cuBLAS (context only): {ms:.3f} ms  {(n*k*2+n*m*2)/ms/1e6:.0f} GB/s")
