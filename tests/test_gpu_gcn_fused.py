"""K2f (csrc/gcn_fused.cu): CSR segment-sum fused with the layer's Linear on tcgen05 (GCNConv forward, GINConv + first Linear of
its MLP), through the C ABI (b2g_segw_gemm): arbitrary CSRs (empty rows, rows of 1..8 and > 8 entries, partial tiles, ghost
sources beyond the target rows), every combination of column / row scale, self term, bias and ReLU against an fp64 torch
reference of the same contraction; the layers fused vs unfused vs the fp64 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _random_csr(N, n_src, seed, max_len=9):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, max_len + 1, size=N)
    deg[rng.integers(0, N, size=max(N // 10, 1))] = 0
    if N > 40:
        deg[3] = 45
        deg[N - 1] = 12
    rowptr = np.zeros(N + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rowptr[-1]))
    return torch.from_numpy(rowptr).int().cuda(), torch.from_numpy(col).int().cuda()


def _ref(x, rowptr, col, N, w, bias, cs, rs, self_coef, relu):
    deg = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(N, device=x.device), deg)
    xs = x.double()[col.long()]
    if cs is not None:
        xs = xs * cs.double()[col.long()].unsqueeze(1)
    z = torch.zeros(N, x.shape[1], dtype=torch.float64, device=x.device).index_add_(0, rows, xs)
    if self_coef:
        z = z + self_coef * x.double()[:N]
    out = z @ w.double().t()
    if rs is not None:
        out = out * rs.double().unsqueeze(1)
    if bias is not None:
        out = out + bias.double()
    return out.clamp_min(0) if relu else out


@pytest.mark.parametrize("C", [64, 256])
@pytest.mark.parametrize("N", [1, 127, 129, 1000, 4100])
@pytest.mark.parametrize("variant", ["gcn", "gin", "plain"])
def test_segw_gemm_kernel_vs_fp64(N, C, variant):
    from gnn_bfs_rans_b200 import ops
    F = 256
    n_src = N + 29
    rowptr, col = _random_csr(N, n_src, seed=N * 3 + C)
    torch.manual_seed(N + C)
    x = torch.randn(n_src, F, device="cuda").bfloat16()
    w = (torch.randn(C, F, device="cuda") / 8).bfloat16()
    bias = torch.randn(C, device="cuda")
    cs = torch.rand(n_src, device="cuda") + 0.2
    rs = torch.rand(N, device="cuda") + 0.2
    kw = {"gcn": dict(col_scale=cs, row_scale=rs, self_coef=0.0, relu=False),
          "gin": dict(col_scale=None, row_scale=None, self_coef=1.25, relu=True),
          "plain": dict(col_scale=None, row_scale=None, self_coef=0.0, relu=False)}[variant]
    b = None if variant == "plain" else bias
    assert ops.segw_gemm_supported(N, F, C, torch.bfloat16)
    out = torch.full((N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.segw_gemm(x, rowptr, col, N, w, b, out=out, **kw)
    ref = _ref(x, rowptr, col, N, w, b, kw["col_scale"], kw["row_scale"], kw["self_coef"], kw["relu"])
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < 2e-2, err
    out2 = torch.empty_like(out)
    ops.segw_gemm(x, rowptr, col, N, w, b, out=out2, **kw)
    assert torch.equal(out, out2)                                            # deterministic
    if variant == "plain":
        empty = rowptr[1:] == rowptr[:-1]
        assert float(out[empty].abs().max()) == 0.0 if bool(empty.any()) else True


def test_gcnconv_and_ginconv_fused_equal_unfused_and_oracle(monkeypatch):
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from oracle import layers_oracle as lo
    nx, ny, nz = 40, 30, 70                                                  # band large enough for the panel order
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device="cuda")
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    ei = torch.cat([ei, torch.stack([torch.randint(0, N, (40,), device="cuda"), torch.full((40,), 11, device="cuda")])], 1)
    torch.manual_seed(0)
    x = torch.randn(N, 256, device="cuda").bfloat16()
    rows = torch.randint(0, N, (3000,))
    rows[:3] = torch.tensor([0, 11, N - 1])
    from oracle import sampled
    nodes, ei_sub, pos = sampled.closure_subgraph(ei.cpu(), rows, N)
    for kind in ("GCN", "GIN"):
        torch.manual_seed(1)
        layer = (b2g.nn.GCNConv(256, 256) if kind == "GCN" else
                 b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256))))
        with torch.no_grad():
            for p_ in layer.parameters():
                if p_.dim() == 1:
                    p_.uniform_(-0.5, 0.5)
        layer = layer.cuda().bfloat16().eval()
        outs, launches = {}, {}
        from gnn_bfs_rans_b200 import _lib
        for path in ("fused", "unfused"):
            monkeypatch.setenv("B2G_GCN_PATH", path)
            monkeypatch.setenv("B2G_GIN_PATH", path)
            with torch.no_grad():
                layer(x, ei)                                                 # graph build + folds outside the count
                c0 = _lib.launch_count()
                outs[path] = layer(x, ei)
                launches[path] = _lib.launch_count() - c0
        assert launches["fused"] < launches["unfused"], (kind, launches)    # the one-kernel path really ran
        ref = sampled.layer_rows(kind, layer.state_dict(), x[nodes.cuda()], ei_sub, pos)
        for path in ("fused", "unfused"):
            got = outs[path][rows.cuda()].double().cpu()
            assert float((got - ref).abs().max() / ref.abs().max()) < 2e-2, (kind, path)
        d = float((outs["fused"].float() - outs["unfused"].float()).abs().max() / outs["unfused"].float().abs().max())
        # GCN: different rounding points (the unfused path rounds the aggregate OR the projection to bf16 in HBM); GIN: the
        # two paths round at the same points and accumulate in the same order, so they may agree to the bit
        assert d < 1.5e-2 and (d > 0 or kind == "GIN"), (kind, d)
    # training-mode GCNConv: fused forward, the usual backward
    monkeypatch.setenv("B2G_GCN_PATH", "fused")
    layer = b2g.nn.GCNConv(256, 256).cuda().bfloat16()
    xg = x[:20000].clone().requires_grad_(True)
    sub = ei[:, (ei[0] < 20000) & (ei[1] < 20000)]
    layer(xg, sub).float().square().mean().backward()
    g1 = xg.grad.clone()
    monkeypatch.setenv("B2G_GCN_PATH", "unfused")
    xg2 = x[:20000].clone().requires_grad_(True)
    layer.zero_grad(set_to_none=True)
    layer(xg2, sub).float().square().mean().backward()
    assert float((g1.float() - xg2.grad.float()).norm() / xg2.grad.float().norm()) < 2e-2
