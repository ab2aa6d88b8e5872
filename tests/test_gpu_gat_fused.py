"""K4f (csrc/gat_fused.cu): the fused attention-weighted aggregation + output projection of GATConv(heads=4, concat=False)
and its attention-weight kernel, through the C ABI (b2g_gat_alpha / b2g_gatw_gemm).
  * kernel level: arbitrary CSRs (empty rows, rows of 1..8, 9..32 and > 32 entries, partial tiles, C = 64 / 128 / 256, the
    perm indirection of the transposed use) against an fp64 torch reference of the same contraction;
  * alpha against the fp64 segment softmax PyG defines (oracle/layers_oracle.py segment_softmax);
  * layer level: GATConv forward + gradients vs the fp64 oracle on meshes large enough for the panel row order, and the
    fused path against the unfused one (same arithmetic, z rounded to bf16 in both) incl. attention dropout (same mask)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H = 4


def _random_csr(N, n_src, seed, max_len=12, hubs=True):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, max_len + 1, size=N)
    deg[rng.integers(0, N, size=max(N // 10, 1))] = 0                      # empty rows
    if hubs and N > 40:
        deg[3] = 45
        deg[N // 2] = 33
        deg[N - 1] = 9
    rowptr = np.zeros(N + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rowptr[-1]))
    return torch.from_numpy(rowptr).int().cuda(), torch.from_numpy(col).int().cuda(), torch.from_numpy(deg).cuda()


def _ref_out(x, rowptr, col, alpha_of_pos, wc, bias, N, F):
    """fp64: out = z Wc^T + b with z_i = [sum_p alpha[p, h] x[col_p]]_h."""
    deg = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(N, device=x.device), deg)
    xs = x.double()[col.long()]                                             # [nnz, F]
    z = torch.zeros(N, H, F, dtype=torch.float64, device=x.device)
    z.index_add_(0, rows, alpha_of_pos.double().unsqueeze(-1) * xs.unsqueeze(1))
    out = z.reshape(N, H * F) @ wc.double().t()
    return out + bias.double() if bias is not None else out


@pytest.mark.parametrize("C", [64, 128, 256])
@pytest.mark.parametrize("N", [1, 5, 127, 128, 129, 1000, 4100])
def test_gatw_gemm_kernel_vs_fp64(N, C):
    from gnn_bfs_rans_b200 import ops
    F = 256
    n_src = N + 37                                                          # gathered rows may lie beyond the target rows (ghosts)
    rowptr, col, _ = _random_csr(N, n_src, seed=N * 7 + C)
    nnz = col.numel()
    torch.manual_seed(N + C)
    x = torch.randn(n_src, F, device="cuda").bfloat16()
    alpha = torch.rand(max(nnz, 1), H, device="cuda")
    wc = (torch.randn(C, H * F, device="cuda") / 16).bfloat16()
    bias = torch.randn(C, device="cuda")
    wp = wc.view(C, H, F // 64, 64).permute(0, 2, 1, 3).reshape(C, H * F).contiguous()
    assert ops.gatw_gemm_supported(N, H, F, C, torch.bfloat16)
    out = torch.full((N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gatw_gemm(x, rowptr, col, None, alpha, wp, bias, N, H, out=out)
    ref = _ref_out(x, rowptr, col, alpha[:nnz], wc, bias, N, F)
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < 2e-2, err
    # rows without entries are exactly the bias
    empty = (rowptr[1:] == rowptr[:-1])
    if bool(empty.any()):
        assert torch.equal(out[empty], bias.bfloat16().expand(int(empty.sum()), C))
    # deterministic: a second launch gives the same bits
    out2 = torch.empty_like(out)
    ops.gatw_gemm(x, rowptr, col, None, alpha, wp, bias, N, H, out=out2)
    assert torch.equal(out, out2)
    # perm indirection (transposed use): alpha stored in another order
    if nnz:
        perm = torch.randperm(nnz, device="cuda").int()
        alpha_p = torch.empty_like(alpha)
        alpha_p[perm.long()] = alpha[:nnz]
        out3 = torch.empty_like(out)
        ops.gatw_gemm(x, rowptr, col, perm, alpha_p, wp, None, N, H, out=out3)
        ref3 = _ref_out(x, rowptr, col, alpha[:nnz], wc, None, N, F)
        assert float((out3.double() - ref3).abs().max() / ref3.abs().max().clamp_min(1e-30)) < 2e-2


def test_gatw_gemm_panel_order_and_band():
    """A band-structured mesh large enough for the panel row order (band >= 4 panels): every row computed exactly once."""
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.graph import Graph
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    nx, ny, nz = 40, 30, 70                                                 # 84 000 rows, band = 1200 ... use the z-stride
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    g = Graph(ei, N)
    csr = g.csr("sl", False)
    torch.manual_seed(0)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    alpha = torch.rand(csr.nnz, H, device="cuda")
    wc = (torch.randn(256, 1024, device="cuda") / 16).bfloat16()
    wp = wc.view(256, H, 4, 64).permute(0, 2, 1, 3).reshape(256, 1024).contiguous()
    outs = []
    for band in (0, g.band(), 40000):
        out = torch.full((N, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gatw_gemm(x, csr.rowptr, csr.col, None, alpha, wp, None, N, H, band=band, out=out)
        assert bool(torch.isfinite(out.float()).all())
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    rows = torch.randint(0, N, (3000,), device="cuda")
    ref = _ref_out(x, csr.rowptr, csr.col, alpha, wc, None, N, 256)[rows]
    assert float((outs[0][rows].double() - ref).abs().max() / ref.abs().max()) < 2e-2


@pytest.mark.parametrize("p_drop", [0.0, 0.3])
def test_gat_alpha_is_pygs_segment_softmax(p_drop):
    from gnn_bfs_rans_b200 import ops
    from oracle import layers_oracle as lo
    N = 3000
    rowptr, col, deg = _random_csr(N, N, seed=11, max_len=10)
    nnz = col.numel()
    torch.manual_seed(3)
    a = torch.randn(N, 2 * H, device="cuda") * 2
    alpha, smax, ssum = ops.gat_alpha(a, rowptr, col, H, 0.2, p_drop, 1234, True)
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), deg.long())
    s = torch.nn.functional.leaky_relu(a.double()[col.long(), :H] + a.double()[rows, H:], 0.2)
    ref = lo.segment_softmax(s.cpu(), rows.cpu(), N)
    got = alpha[:nnz].double().cpu()
    if p_drop == 0.0:
        assert float((got - ref).abs().max()) < 1e-5
    else:
        keep = got != 0
        frac = float(keep.double().mean())
        assert abs(frac - (1 - p_drop)) < 0.02, frac
        assert float((got[keep] * (1 - p_drop) - ref[keep]).abs().max()) < 1e-5      # kept entries scaled by 1/(1-p)
        again, _, _ = ops.gat_alpha(a, rowptr, col, H, 0.2, p_drop, 1234, False)
        assert torch.equal(again, alpha)                                               # same seed -> same mask
        other, _, _ = ops.gat_alpha(a, rowptr, col, H, 0.2, p_drop, 99, False)
        assert not torch.equal(other, alpha)
    # statistics the backward kernels read: max of the scores and the sum of exp (+ 1e-16) per (row, head)
    nz = deg > 0
    mref = torch.zeros(N, H, dtype=torch.float64).scatter_reduce_(0, rows.cpu().view(-1, 1).expand(-1, H), s.cpu(), "amax", include_self=False)
    assert float((smax.double().cpu() - mref)[nz.cpu()].abs().max()) < 1e-5
    zref = torch.zeros(N, H, dtype=torch.float64).index_add_(0, rows.cpu(), (s.cpu() - mref[rows.cpu()]).exp())
    assert float(((ssum.double().cpu() - zref) / zref.clamp_min(1e-30))[nz.cpu()].abs().max()) < 1e-5


def _layer(C=256):
    import gnn_bfs_rans_b200 as b2g
    torch.manual_seed(1234)
    m = b2g.nn.GATConv(256, C, heads=4, concat=False, dropout=0.2)
    with torch.no_grad():
        m.bias.uniform_(-0.5, 0.5)
    return m.cuda().bfloat16()


@pytest.mark.parametrize("C", [128, 256])
def test_gatconv_fused_equals_unfused_and_oracle(C, monkeypatch):
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from oracle import layers_oracle as lo
    nx, ny, nz = 24, 20, 18
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    extra = torch.stack([torch.randint(0, N, (60,), device="cuda"), torch.full((60,), 17, device="cuda")])   # a hub target
    ei = torch.cat([ei, extra], 1)
    m = _layer(C).eval()
    torch.manual_seed(5)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    gout = torch.randn(N, C, device="cuda").bfloat16()
    res = {}
    for path in ("fused", "unfused"):
        monkeypatch.setenv("B2G_GAT_PATH", "" if path == "fused" else "unfused")
        xg = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        out = m(xg, ei)
        out.backward(gout)
        res[path] = (out.detach(), xg.grad, {k: p.grad.clone() for k, p in m.named_parameters()})
    p = {k: v.detach().double().cpu().requires_grad_(True) for k, v in m.state_dict().items()}
    x64 = x.double().cpu().requires_grad_(True)
    ref = lo.gat_conv(x64, ei.cpu(), p["lin.weight"], p["att_src"], p["att_dst"], p["bias"], heads=4)
    ref.backward(gout.double().cpu())
    rel = lambda a_, b_: float((a_.double().cpu() - b_).abs().max() / b_.abs().max().clamp_min(1e-30))
    assert rel(res["fused"][0], ref.detach()) < 2e-2
    assert rel(res["unfused"][0], ref.detach()) < 2e-2
    assert rel(res["fused"][0], res["unfused"][0].double().cpu()) < 1.2e-2      # two bf16 roundings of the same value
    for path in ("fused", "unfused"):                                            # same backward kernels either way
        assert float((res[path][1].double().cpu() - x64.grad).norm() / x64.grad.norm()) < 4e-2, path
        for k in ("lin.weight", "bias"):
            assert float((res[path][2][k].double().cpu() - p[k].grad).norm() / p[k].grad.norm()) < 4e-2, (path, k)


def test_gatconv_fused_dropout_uses_the_mask_the_backward_regenerates(monkeypatch):
    """Training mode: the fused forward draws its attention-dropout mask in b2g_gat_alpha with the (seed, position) key the
    backward kernels regenerate it from; under the same torch seed fused and unfused runs see the same mask."""
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    nx, ny, nz = 16, 14, 12
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    m = _layer().train()
    x = torch.randn(N, 256, device="cuda").bfloat16()
    res = {}
    for path in ("fused", "unfused"):
        monkeypatch.setenv("B2G_GAT_PATH", "" if path == "fused" else "unfused")
        torch.manual_seed(77)
        xg = x.clone().requires_grad_(True)
        out = m(xg, ei)
        out.float().square().mean().backward()
        res[path] = (out.detach().float(), xg.grad.float())
    scale = res["unfused"][0].abs().max()
    assert float((res["fused"][0] - res["unfused"][0]).abs().max() / scale) < 1.2e-2
    assert float((res["fused"][1] - res["unfused"][1]).norm() / res["unfused"][1].norm()) < 2e-2
    m.eval()
    with torch.no_grad():
        ev = m(x, ei).float()
    assert float((res["fused"][0] - ev).abs().max() / scale) > 5e-2           # the mask is really applied


@pytest.mark.parametrize("C", [128, 256])
@pytest.mark.parametrize("train", [False, True])
def test_transformerconv_fused_equals_unfused_and_oracle(C, train, monkeypatch):
    """Fused TransformerConv forward (b2g_tz_alpha + b2g_gatw_gemm_ex: value projection, s.bv and skip terms in one kernel, no
    z_aug) against the unfused aggregate-first path and the fp64 oracle; gradients of both against the oracle (the fused
    path derives d W_out from y instead of z_aug); with attention dropout both paths draw the same mask."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from oracle import layers_oracle as lo
    nx, ny, nz = 24, 20, 18
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    extra = torch.stack([torch.randint(0, N, (60,), device="cuda"), torch.full((60,), 17, device="cuda")])   # a hub target
    ei = torch.cat([ei, extra], 1)
    ei = ei[:, ei[1] != 5]                                                  # node 5 has no incoming edge at all
    torch.manual_seed(1234)
    m = b2g.nn.TransformerConv(256, C, heads=4, concat=False, dropout=0.2)
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() == 1:
                p_.uniform_(-0.5, 0.5)
    m = m.cuda().bfloat16().train(train)
    torch.manual_seed(5)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    gout = torch.randn(N, C, device="cuda").bfloat16()
    res = {}
    for path in ("fused", "unfused"):
        monkeypatch.setenv("B2G_TCONV_PATH", "" if path == "fused" else "unfused")
        torch.manual_seed(77)
        xg = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        out = m(xg, ei)
        out.backward(gout)
        res[path] = (out.detach(), xg.grad, {k: p.grad.clone() for k, p in m.named_parameters()})
    rel = lambda a_, b_: float((a_.double().cpu() - b_.double().cpu()).abs().max() / b_.double().abs().max().clamp_min(1e-30))
    l2 = lambda a_, b_: float((a_.double().cpu() - b_.double().cpu()).norm() / b_.double().norm().clamp_min(1e-30))
    assert rel(res["fused"][0], res["unfused"][0]) < 1.2e-2
    assert l2(res["fused"][1], res["unfused"][1]) < 2e-2
    for k in res["unfused"][2]:
        if float(res["unfused"][2][k].abs().max()) > 0:
            assert l2(res["fused"][2][k], res["unfused"][2][k]) < 3e-2 or k == "lin_key.bias", k
    if not train:
        p = {k: v.detach().double().cpu().requires_grad_(True) for k, v in m.state_dict().items()}
        x64 = x.double().cpu().requires_grad_(True)
        ref = lo.transformer_conv(x64, ei.cpu(), p["lin_query.weight"], p["lin_query.bias"], p["lin_key.weight"], p["lin_key.bias"],
                                  p["lin_value.weight"], p["lin_value.bias"], p["lin_skip.weight"], p["lin_skip.bias"], heads=4)
        ref.backward(gout.double().cpu())
        for path in ("fused", "unfused"):
            assert rel(res[path][0], ref.detach()) < 2e-2, path
            assert l2(res[path][1], x64.grad) < 2e-2, path
            for k in ("lin_value.weight", "lin_value.bias", "lin_skip.weight", "lin_query.weight"):
                assert l2(res[path][2][k], p[k].grad) < 2e-2, (path, k)


@pytest.mark.parametrize("N", [1, 37, 1000, 5000])
@pytest.mark.parametrize("p_drop", [0.0, 0.3])
@pytest.mark.parametrize("with_bias", [False, True])
def test_tz_alpha_tensor_core_logits_vs_simt_and_fp64(N, p_drop, with_bias):
    """b2g_tz_alpha, bf16 F = 256: the m16n8k16 tile-product kernel (impl 0) against the SIMT dot products (impl 1) and an fp64
    softmax of u_ih . x_j over random CSRs with empty rows, rows of 1..8, 9..31 and > 32 entries; same dropout mask."""
    from gnn_bfs_rans_b200 import ops
    import numpy as np
    rng = np.random.default_rng(N + 7)
    deg = rng.integers(0, 10, size=N)
    if N > 30:
        deg[4] = 70
        deg[9] = 21
        deg[N - 1] = 33
        deg[7] = 0
    rowptr = np.zeros(N + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    n_src = N + 11
    col = torch.from_numpy(rng.integers(0, n_src, size=max(nnz, 1))).int().cuda()[:nnz] if nnz else torch.zeros(0, dtype=torch.int32, device="cuda")
    rp = torch.from_numpy(rowptr).int().cuda()
    torch.manual_seed(N)
    x = torch.randn(n_src, 256, device="cuda").bfloat16()
    u = (torch.randn(N, 1024, device="cuda") / 16).bfloat16()
    eb = torch.randn(max(nnz, 1), 4, device="cuda") if with_bias else None
    res = {}
    for impl in (0, 1):
        res[impl] = ops.tz_alpha(x, u, 4, rp, col, p_drop, 99, band=0, edge_bias=eb, impl=impl)
    if nnz:
        rows = torch.repeat_interleave(torch.arange(N, device="cuda"), torch.from_numpy(deg).cuda())
        logit = (u.double().view(N, 4, 256)[rows] * x.double()[col.long()].unsqueeze(1)).sum(-1)
        if with_bias:
            logit = logit + eb[:nnz].double()
        mx = torch.full((N, 4), -float("inf"), dtype=torch.float64, device="cuda").scatter_reduce(0, rows[:, None].expand(-1, 4), logit, "amax")
        ex = (logit - mx[rows]).exp()
        den = torch.zeros(N, 4, dtype=torch.float64, device="cuda").index_add_(0, rows, ex)
        ref = ex / (den[rows] + 1e-16)
        for impl in (0, 1):
            assert float((res[impl][0][:nnz].double() - ref).abs().max()) < 2e-5, impl
        assert float((res[0][0][:nnz] - res[1][0][:nnz]).abs().max()) < 2e-6         # same products, another summation order
        if p_drop > 0:
            keep0, keep1 = res[0][1][:nnz] != 0, res[1][1][:nnz] != 0
            assert torch.equal(keep0 | (res[0][0][:nnz] == 0), keep1 | (res[1][0][:nnz] == 0))   # the same Philox mask
            assert float((res[0][1][:nnz] - res[1][1][:nnz]).abs().max()) < 4e-6
            post = res[0][1][:nnz]
        else:
            assert res[0][1] is None
            post = res[0][0][:nnz]
        ssum = torch.zeros(N, 4, device="cuda").index_add_(0, rows, post)
        assert float((res[0][2] - ssum).abs().max()) < 1e-5
    empty = torch.from_numpy(deg == 0).cuda()
    if bool(empty.any()):
        assert float(res[0][2][empty].abs().max()) == 0.0


@pytest.mark.parametrize("N", [1, 37, 1000, 5000])
@pytest.mark.parametrize("use_perm", [False, True])
def test_weighted_row_sums_tensor_core_vs_fp64(N, use_perm):
    """b2g_gatz_bwd_src for bf16 rows of 512 bytes (gatz_bwd_src_mma_kernel: the 4-head weighted sums as m16n8k8 tile products,
    weights split into bf16 head + remainder) against fp64 on random CSRs with empty rows and rows of 1..8, 9..31, > 32 entries;
    the companion row sums (d a_src)."""
    from gnn_bfs_rans_b200 import ops
    import numpy as np
    rng = np.random.default_rng(N + 3)
    deg = rng.integers(0, 10, size=N)
    if N > 30:
        deg[4] = 70
        deg[9] = 21
        deg[N - 1] = 33
        deg[7] = 0
    rowptr = np.zeros(N + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    n_src = N + 13
    col = torch.from_numpy(rng.integers(0, n_src, size=max(nnz, 1))).int().cuda()
    perm = torch.from_numpy(rng.permutation(max(nnz, 1))).int().cuda() if use_perm else None
    rp = torch.from_numpy(rowptr).int().cuda()
    torch.manual_seed(N)
    x = torch.randn(n_src, 256, device="cuda").bfloat16()
    w = torch.randn(max(nnz, 1), 4, device="cuda")
    out = torch.full((N, 1024), float("nan"), device="cuda", dtype=torch.bfloat16)
    d_a = torch.full((N, 8), float("nan"), device="cuda")
    ops.seg_wsum4(x, w, rp, col, perm, out, d_a=d_a)
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), torch.from_numpy(deg).cuda())
    wp = w[perm.long()[:nnz]] if use_perm else w[:nnz]
    ref = torch.zeros(N, 4, 256, dtype=torch.float64, device="cuda")
    if nnz:
        ref.index_add_(0, rows, wp.double().unsqueeze(2) * x.double()[col.long()[:nnz]].unsqueeze(1))
    err = float((out.double().view(N, 4, 256) - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < 6e-3, err                                                   # bf16 rounding of the output only
    sums = torch.zeros(N, 4, dtype=torch.float64, device="cuda")
    if nnz:
        sums.index_add_(0, rows, wp.double())
    assert float((d_a[:, :4].double() - sums).abs().max()) < 1e-4
    # the weights keep fp32 accuracy: the error is the bf16 rounding of the result (2^-8 |y|) plus 1e-5 of sum |w| |x| — weights
    # rounded to bf16 would show 2^-9 of that sum
    mag = torch.zeros(N, 4, 256, dtype=torch.float64, device="cuda")
    if nnz:
        mag.index_add_(0, rows, wp.double().abs().unsqueeze(2) * x.double()[col.long()[:nnz]].abs().unsqueeze(1))
    excess = (out.double().view(N, 4, 256) - ref).abs() - (2.0 ** -8) * ref.abs() - 1e-5 * mag
    assert float(excess.max()) <= 0.0


@pytest.mark.parametrize("N", [1, 7, 8, 1000, 4099])
def test_rowdot8_tensor_core_vs_fp64(N):
    """b2g_rowdot8 for bf16 rows of 512 bytes (rowdot8_mma_kernel: V as bf16 head + remainder B fragments) against fp64: the fp32
    vectors keep their accuracy (a bf16-rounded V would show 2^-9)."""
    from gnn_bfs_rans_b200 import ops
    torch.manual_seed(N)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    v = torch.randn(8, 256, device="cuda")
    a = ops.rowdot8(x, v)
    ref = x.double() @ v.double().T
    mag = x.double().abs() @ v.double().abs().T
    assert a.shape == (N, 8)
    assert float(((a.double() - ref).abs() / mag).max()) < 2e-5


@pytest.mark.parametrize("edge", [False, True])
@pytest.mark.parametrize("train", [False, True])
def test_gatconv_softmax_inside_the_fused_kernel_equals_separate_kernel(train, edge, monkeypatch):
    """Rows of <= 8 entries: b2g_gatw_gemm_sm (softmax in the gather warps' per-tile prologue) against b2g_gat_alpha +
    b2g_gatw_gemm: same formulas, the 8-term sums in another order -> outputs equal to bf16 rounding, the statistics for the
    backward pass to fp32 rounding, gradients alike; with dropout both draw the same mask."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200.graph import graph_of
    nx, ny, nz = 30, 20, 25
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    assert graph_of(ei, N).max_degree("sl") <= 8
    torch.manual_seed(3)
    m = b2g.nn.GATConv(256, 256, heads=4, concat=False, dropout=0.3, edge_dim=4 if edge else None).cuda().bfloat16().train(train)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    ea = torch.randn(ei.shape[1], 4, device="cuda").bfloat16() if edge else None      # the edge term of the logits (edge_bias)
    gout = torch.randn(N, 256, device="cuda").bfloat16()
    res = {}
    for mode in ("", "separate"):
        monkeypatch.setenv("B2G_GAT_SOFTMAX", mode)
        torch.manual_seed(11)
        xg = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        out = m(xg, ei, edge_attr=ea) if edge else m(xg, ei)
        out.backward(gout)
        res[mode] = (out.detach().float(), xg.grad.float(), m.att_src.grad.float().clone())
    scale = res["separate"][0].abs().max()
    assert float((res[""][0] - res["separate"][0]).abs().max() / scale) < 8e-3          # one bf16 ulp at most
    assert float((res[""][0] != res["separate"][0]).float().mean()) < 1e-2              # ... and on few elements
    assert float((res[""][1] - res["separate"][1]).norm() / res["separate"][1].norm()) < 5e-3
    assert float((res[""][2] - res["separate"][2]).norm() / res["separate"][2].norm()) < 5e-3


def test_transformerconv_fused_eval_without_empty_rows_folds_the_value_bias(monkeypatch):
    """No attention dropout and every row non-empty: s_ih = 1, so sum_h s_ih bv_h / H joins the bias of the skip GEMM and the fused
    kernel runs without the per-row s.bv term; against the unfused path and the fp64 oracle, with large value biases."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.graph import graph_of
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from oracle import layers_oracle as lo
    nx, ny, nz = 22, 18, 16
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    assert graph_of(ei, N).min_degree("raw") >= 1
    torch.manual_seed(21)
    m = b2g.nn.TransformerConv(256, 256, heads=4, concat=False)
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() == 1:
                p_.uniform_(-2.0, 2.0)
    m = m.cuda().bfloat16().eval()
    x = torch.randn(N, 256, device="cuda").bfloat16()
    outs = {}
    for path in ("", "unfused"):
        monkeypatch.setenv("B2G_TCONV_PATH", path)
        with torch.no_grad():
            outs[path] = m(x, ei).double().cpu()
    p = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    ref = lo.transformer_conv(x.double().cpu(), ei.cpu(), p["lin_query.weight"], p["lin_query.bias"], p["lin_key.weight"],
                              p["lin_key.bias"], p["lin_value.weight"], p["lin_value.bias"], p["lin_skip.weight"], p["lin_skip.bias"],
                              heads=4)
    scale = ref.abs().max()
    for path in ("", "unfused"):
        assert float((outs[path] - ref).abs().max() / scale) < 2e-2, path
    assert float((outs[""] - outs["unfused"]).abs().max() / scale) < 1.2e-2
