"""K6 parity: the per-node Linear (fwd / dgrad / wgrad, bias / ReLU / aux-split epilogues) vs a plain
torch fp64 reference of the same op.  Tolerances: 1e-5 (fp32, exact-fp32 SIMT path) and 2e-2 (bf16),
relative to the tensor's max-abs (BASELINE.json north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}
IMPLS = [1, 0]   # SIMT forced, then auto (tcgen05 where covered)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,k,m", [(1000, 128, 128), (4099, 256, 256), (777, 256, 1024), (130, 64, 40), (5, 8, 8),
                                   (3000, 256, 3328), (2000, 1032, 256), (1500, 3336, 256), (900, 520, 128)])
def test_linear_fwd_bwd(impl, dtype, n, k, m):
    from gnn_bfs_rans_b200 import ops
    ops.GEMM_IMPL = impl
    try:
        torch.manual_seed(n + k + m)
        x = torch.randn(n, k, device='cuda').to(dtype)
        w = (torch.randn(m, k, device='cuda') / k ** 0.5).to(dtype)
        b = torch.randn(m, device='cuda')
        xd, wd, bd = x.double(), w.double(), b.double()
        y, _ = ops.linear_fwd(x, w, b, act=0)
        assert rel(y, xd @ wd.T + bd) < TOL[dtype]
        y, _ = ops.linear_fwd(x, w, b, act=1)
        assert rel(y, torch.relu(xd @ wd.T + bd)) < TOL[dtype]
        rs = torch.rand(n, device='cuda') + 0.5
        y, _ = ops.linear_fwd(x, w, None, row_scale=rs)
        assert rel(y, rs.double()[:, None] * (xd @ wd.T)) < TOL[dtype]
        gy = torch.randn(n, m, device='cuda').to(dtype)
        assert rel(ops.linear_dgrad(gy, w), gy.double() @ wd) < TOL[dtype]
        dw, db = ops.linear_wgrad(gy, x)
        assert dw.dtype == torch.float32
        assert rel(dw, gy.double().T @ xd) < (1e-5 if dtype == torch.float32 else 1e-4)
        assert rel(db, gy.double().sum(0)) < 1e-5
    finally:
        ops.GEMM_IMPL = 0


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_aux_split(impl, dtype):
    """GAT: the 2H attention-logit columns ride along and land in an fp32 aux matrix."""
    from gnn_bfs_rans_b200 import ops
    ops.GEMM_IMPL = impl
    try:
        n, k, hc, extra = 2500, 256, 1024, 8
        x = torch.randn(n, k, device='cuda').to(dtype)
        w = (torch.randn(hc + extra, k, device='cuda') / 16).to(dtype)
        y, aux = ops.linear_fwd(x, w, None, m_main=hc)
        ref = x.double() @ w.double().T
        assert y.shape == (n, hc) and aux.shape == (n, extra) and aux.dtype == torch.float32
        assert rel(y, ref[:, :hc]) < TOL[dtype]
        assert rel(aux, ref[:, hc:]) < (1e-5 if dtype == torch.float32 else 1e-5)   # aux stays fp32-accurate
    finally:
        ops.GEMM_IMPL = 0


def test_strided_views_and_autograd():
    from gnn_bfs_rans_b200 import functional as Fn
    torch.manual_seed(0)
    big = torch.randn(600, 512, device='cuda')
    x = big[:, 128:384].requires_grad_(True)       # row stride 512, 16-byte aligned offset
    lin = torch.nn.Linear(256, 64).cuda()
    y = Fn.linear(x, lin.weight, lin.bias, act=1)
    y.square().sum().backward()
    gx, gw, gb = x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()
    x.grad = None
    lin.zero_grad()
    yr = torch.relu(torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double()))
    yr.square().sum().backward()
    assert rel(y, yr) < 1e-5 and rel(gx, x.grad) < 1e-5
    assert rel(gw, lin.weight.grad) < 1e-5 and rel(gb, lin.bias.grad) < 1e-5


def test_colsum_and_rows():
    from gnn_bfs_rans_b200 import ops
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(10007, 256, device='cuda').to(dtype)
        assert rel(ops.colsum(x), x.double().sum(0)) < 1e-5
        idx = torch.randperm(10007, device='cuda')[:3000].int()
        assert torch.equal(ops.rows_gather(x, idx), x[idx.long()])
        wide = torch.zeros(10007, 640, device='cuda', dtype=dtype)               # idx None: a copy into a column block
        ops.rows_gather(x, None, out=wide[:, 128:384])
        assert torch.equal(wide[:, 128:384], x) and float(wide[:, :128].abs().max()) == 0 and float(wide[:, 384:].abs().max()) == 0
        y = x.clone()
        add = torch.randn(3000, 256, device='cuda').to(dtype)
        ops.rows_scatter_add(y, idx, add)
        ref = x.clone()
        ref[idx.long()] = (x[idx.long()].float() + add.float()).to(dtype)
        assert torch.equal(y, ref)


@pytest.mark.parametrize("n,k,m", [(2 * 148 * 128 + 77, 1032, 256), (2 * 148 * 128 + 128 + 5, 3336, 200), (3 * 148 * 128, 520, 256),
                                   (40000, 256, 1024), (40000, 256, 1040)])
def test_linear_large_row_counts_plans(n, k, m):
    """bf16 tcgen05 Linear at row counts that select the large-problem plans (gemm_tc.cu): two row tiles per W k-block for long
    reductions (plan 2: odd and even numbers of row tiles, a partial last tile), one resident W block per CTA group for
    k <= 256 and wide outputs (plan 1, incl. a ragged last column tile), TMA tile stores; every output row against fp64."""
    from gnn_bfs_rans_b200 import ops
    torch.manual_seed(k + m)
    x = torch.randn(n, k, device='cuda').bfloat16()
    w = (torch.randn(m, k, device='cuda') / k ** 0.5).bfloat16()
    b = torch.randn(m, device='cuda')
    rs = torch.rand(n, device='cuda') + 0.5
    y = torch.full((n, m), float('nan'), device='cuda', dtype=torch.bfloat16)
    ops.linear_fwd(x, w, b, row_scale=rs, act=1, out=y)
    for r0 in range(0, n, 8192):
        xs = x[r0:r0 + 8192].double()
        ref = torch.relu(rs[r0:r0 + 8192].double()[:, None] * (xs @ w.double().T) + b.double())
        assert rel(y[r0:r0 + 8192], ref) < 2e-2, r0
    assert bool(torch.isfinite(y.float()).all())


@pytest.mark.parametrize("n,k,m", [(4099, 256, 256), (700, 256, 128), (40000, 1024, 256), (130, 64, 40)])
def test_dgrad_with_relu_mask_epilogue(n, k, m):
    """b2g_linear_fwd_masked: dx [n, k] = (dy [n, m] W [m, k]) where mask > 0 — aten.threshold_backward fused into the dgrad GEMM
    (bf16, tcgen05, reductions m <= 256: the resident-W plan)."""
    from gnn_bfs_rans_b200 import ops
    torch.manual_seed(n + m)
    dy = torch.randn(n, m, device='cuda').bfloat16()
    w = (torch.randn(m, k, device='cuda') / m ** 0.5).bfloat16()
    mask = torch.relu(torch.randn(n, k, device='cuda')).bfloat16()
    got = ops.linear_dgrad_masked(dy, w, mask)
    assert got is not None
    ref = (dy.double() @ w.double()) * (mask > 0)
    assert rel(got, ref) < 2e-2
    assert bool((got[mask <= 0] == 0).all())
    assert ops.linear_dgrad_masked(dy.float(), w.float(), mask.float()) is None       # fp32: the caller masks separately
    wide = torch.randn(512, 320, device='cuda').bfloat16()                            # reduction > 256: no fused epilogue either
    assert ops.linear_dgrad_masked(wide, torch.randn(320, 64, device='cuda').bfloat16(), torch.ones(512, 64, device='cuda').bfloat16()) is None


def test_gin_mlp_single_node_backward_equals_two_linears():
    """GINConv's Linear-ReLU-Linear as one autograd node (functional.MLP2Fn) against the two LinearFn nodes + threshold_backward."""
    from gnn_bfs_rans_b200 import functional as Fn
    torch.manual_seed(5)
    n = 5000
    for dtype in (torch.bfloat16, torch.float32):
        x = torch.randn(n, 256, device='cuda').to(dtype)
        ps = [(torch.randn(256, 256, device='cuda') / 16).to(dtype), torch.randn(256, device='cuda').to(dtype),
              (torch.randn(128, 256, device='cuda') / 16).to(dtype), torch.randn(128, device='cuda').to(dtype)]
        gout = torch.randn(n, 128, device='cuda').to(dtype)
        res = []
        for fused in (True, False):
            xs = x.clone().requires_grad_(True)
            pp = [p.clone().requires_grad_(True) for p in ps]
            y = Fn.mlp2(xs, *pp) if fused else Fn.linear(Fn.linear(xs, pp[0], pp[1], act=1), pp[2], pp[3])
            y.backward(gout)
            res.append([y.detach()] + [t.grad for t in [xs] + pp])
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        for a_, b_ in zip(*res):
            assert rel(a_, b_) < tol
