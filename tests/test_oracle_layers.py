"""CPU-only pins of oracle/layers_oracle.py.  PyG is not installable here (PARITY UNPINNED at that
boundary), so the restatement is cross-checked against an independent dense-matrix formulation of
each layer's published definition (Kipf & Welling GCN, Velickovic GAT, Xu GIN, Shi TransformerConv)
on multigraphs with duplicates, self loops and isolated nodes."""
import math

import numpy as np
import pytest
import torch

from oracle import layers_oracle as lo


def graph(N=12, E=60, seed=0):
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, N - 2, size=(2, E))            # last two nodes isolated
    ei[1, :6] = ei[0, :6]                               # self loops
    ei[:, 6:9] = ei[:, 9:12]                            # duplicates
    return torch.from_numpy(ei)


def dense_adj(ei, N, drop_loops):
    A = torch.zeros(N, N, dtype=torch.float64)          # A[i,j] = multiplicity of edge j -> i
    for s, d in ei.t().tolist():
        if drop_loops and s == d:
            continue
        A[d, s] += 1
    return A


def test_gcn_equals_dense_normalised_adjacency():
    N, F = 12, 5
    ei = graph(N)
    x, W, b = torch.randn(N, F, dtype=torch.float64), torch.randn(4, F, dtype=torch.float64), torch.randn(4, dtype=torch.float64)
    A = dense_adj(ei, N, True) + torch.eye(N, dtype=torch.float64)
    dinv = A.sum(1).pow(-0.5)
    ref = (dinv[:, None] * A * dinv[None, :]) @ (x @ W.T) + b
    torch.testing.assert_close(lo.gcn_conv(x, ei, W, b), ref, rtol=1e-12, atol=1e-12)


def test_gin_equals_dense():
    N, F = 12, 5
    ei = graph(N)
    x = torch.randn(N, F, dtype=torch.float64)
    w1, b1, w2, b2 = (torch.randn(F, F, dtype=torch.float64), torch.randn(F, dtype=torch.float64),
                      torch.randn(F, F, dtype=torch.float64), torch.randn(F, dtype=torch.float64))
    h = dense_adj(ei, N, False) @ x + 1.25 * x
    ref = torch.relu(h @ w1.T + b1) @ w2.T + b2
    torch.testing.assert_close(lo.gin_conv(x, ei, lo.gin_mlp(w1, b1, w2, b2), eps=0.25), ref, rtol=1e-12, atol=1e-12)


def test_gat_equals_per_edge_loops():
    N, F, H, C = 12, 6, 4, 3
    ei = graph(N)
    x = torch.randn(N, F, dtype=torch.float64)
    W = torch.randn(H * C, F, dtype=torch.float64)
    a_s, a_d, b = (torch.randn(1, H, C, dtype=torch.float64), torch.randn(1, H, C, dtype=torch.float64),
                   torch.randn(C, dtype=torch.float64))
    xs = (x @ W.T).view(N, H, C)
    edges = [(s, d) for s, d in ei.t().tolist() if s != d] + [(v, v) for v in range(N)]
    ref = torch.zeros(N, H, C, dtype=torch.float64)
    for i in range(N):
        inc = [s for s, d in edges if d == i]
        for h in range(H):
            sc = torch.stack([torch.nn.functional.leaky_relu((xs[j, h] * a_s[0, h]).sum() + (xs[i, h] * a_d[0, h]).sum(), 0.2)
                              for j in inc])
            al = torch.softmax(sc, 0)
            for a, j in zip(al, inc):
                ref[i, h] += a * xs[j, h]
    torch.testing.assert_close(lo.gat_conv(x, ei, W, a_s, a_d, b, heads=H), ref.mean(1) + b, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(lo.gat_conv(x, ei, W, a_s, a_d, None, heads=H, concat=True), ref.reshape(N, H * C), rtol=1e-9, atol=1e-12)


def test_gat_with_edge_features_equals_per_edge_loops():
    """GATConv(edge_dim): loops dropped with their attributes, new loops carry the mean incoming attribute, the edge term
    enters the logits only."""
    N, F, H, C, D = 12, 6, 4, 3, 4
    ei = graph(N)
    E = ei.shape[1]
    x = torch.randn(N, F, dtype=torch.float64)
    ea = torch.randn(E, D, dtype=torch.float64)
    W = torch.randn(H * C, F, dtype=torch.float64)
    a_s, a_d, a_e = (torch.randn(1, H, C, dtype=torch.float64) for _ in range(3))
    we = torch.randn(H * C, D, dtype=torch.float64)
    b = torch.randn(C, dtype=torch.float64)
    xs = (x @ W.T).view(N, H, C)
    edges = [(s, d, ea[e]) for e, (s, d) in enumerate(ei.t().tolist()) if s != d]
    ref = torch.zeros(N, H, C, dtype=torch.float64)
    for i in range(N):
        inc = [(s, at) for s, d, at in edges if d == i]
        mean_attr = torch.stack([at for _, at in inc]).mean(0) if inc else torch.zeros(D, dtype=torch.float64)
        inc = inc + [(i, mean_attr)]
        for h in range(H):
            lg = []
            for j, at in inc:
                emb = (we @ at).view(H, C)
                lg.append((xs[j, h] * a_s[0, h]).sum() + (xs[i, h] * a_d[0, h]).sum() + (emb[h] * a_e[0, h]).sum())
            al = torch.softmax(torch.nn.functional.leaky_relu(torch.stack(lg), 0.2), 0)
            for w, (j, _) in zip(al, inc):
                ref[i, h] += w * xs[j, h]
    out = lo.gat_conv(x, ei, W, a_s, a_d, b, heads=H, edge_attr=ea, we=we, att_edge=a_e)
    torch.testing.assert_close(out, ref.mean(1) + b, rtol=1e-9, atol=1e-12)


def test_transformer_equals_per_edge_loops():
    N, F, H, C = 12, 6, 4, 3
    ei = graph(N)
    x = torch.randn(N, F, dtype=torch.float64)
    mk = lambda o: (torch.randn(o, F, dtype=torch.float64), torch.randn(o, dtype=torch.float64))
    (wq, bq), (wk, bk), (wv, bv), (ws, bs) = mk(H * C), mk(H * C), mk(H * C), mk(C)
    q, k, v = [(x @ w.T + b).view(N, H, C) for w, b in ((wq, bq), (wk, bk), (wv, bv))]
    edges = ei.t().tolist()
    ref = torch.zeros(N, H, C, dtype=torch.float64)
    for i in range(N):
        inc = [s for s, d in edges if d == i]
        if not inc:
            continue
        for h in range(H):
            al = torch.softmax(torch.stack([(q[i, h] * k[j, h]).sum() / math.sqrt(C) for j in inc]), 0)
            for a, j in zip(al, inc):
                ref[i, h] += a * v[j, h]
    out = lo.transformer_conv(x, ei, wq, bq, wk, bk, wv, bv, ws, bs, heads=H)
    torch.testing.assert_close(out, ref.mean(1) + x @ ws.T + bs, rtol=1e-9, atol=1e-12)


def test_transformer_with_edge_features_equals_per_edge_loops():
    """edge_dim (PyG TransformerConv.message): lin_edge(edge_attr) joins the keys and the values of every edge."""
    N, F, H, C, D = 12, 6, 4, 3, 4
    ei = graph(N)
    E = ei.shape[1]
    x = torch.randn(N, F, dtype=torch.float64)
    ea = torch.randn(E, D, dtype=torch.float64)
    mk = lambda o: (torch.randn(o, F, dtype=torch.float64), torch.randn(o, dtype=torch.float64))
    (wq, bq), (wk, bk), (wv, bv), (ws, bs) = mk(H * C), mk(H * C), mk(H * C), mk(C)
    we = torch.randn(H * C, D, dtype=torch.float64)
    q, k, v = [(x @ w.T + b).view(N, H, C) for w, b in ((wq, bq), (wk, bk), (wv, bv))]
    emb = (ea @ we.T).view(E, H, C)
    edges = ei.t().tolist()
    ref = torch.zeros(N, H, C, dtype=torch.float64)
    for i in range(N):
        inc = [(e, s) for e, (s, d) in enumerate(edges) if d == i]
        if not inc:
            continue
        for h in range(H):
            al = torch.softmax(torch.stack([(q[i, h] * (k[j, h] + emb[e, h])).sum() / math.sqrt(C) for e, j in inc]), 0)
            for a, (e, j) in zip(al, inc):
                ref[i, h] += a * (v[j, h] + emb[e, h])
    out = lo.transformer_conv(x, ei, wq, bq, wk, bk, wv, bv, ws, bs, heads=H, edge_attr=ea, we=we)
    torch.testing.assert_close(out, ref.mean(1) + x @ ws.T + bs, rtol=1e-9, atol=1e-12)
    # without an edge_attr the lin_edge weights are unused (PyG applies lin_edge only to a given edge_attr)
    torch.testing.assert_close(lo.transformer_conv(x, ei, wq, bq, wk, bk, wv, bv, ws, bs, heads=H, we=we),
                               lo.transformer_conv(x, ei, wq, bq, wk, bk, wv, bv, ws, bs, heads=H))


def test_transformer_edge_folding_algebra():
    """The aggregate-first form the CUDA path uses (nn.TransformerConv._folded(with_edge=True)), evaluated densely on the
    CPU in fp64: logits u_i . x_j + r_ih . a_ij, output [z | s | x | m] W_out^T + b — equals the oracle."""
    from gnn_bfs_rans_b200.nn import TransformerConv
    torch.manual_seed(5)
    N, F, H, D = 14, 8, 4, 4
    ei = graph(N)
    E = ei.shape[1]
    layer = TransformerConv(F, F, heads=H, concat=False, edge_dim=D).double()
    assert list(layer.state_dict()) == ['lin_key.weight', 'lin_key.bias', 'lin_query.weight', 'lin_query.bias',
                                        'lin_value.weight', 'lin_value.bias', 'lin_edge.weight', 'lin_skip.weight',
                                        'lin_skip.bias']
    assert layer.lin_edge.weight.shape == (H * F, D)
    x = torch.randn(N, F, dtype=torch.float64)
    ea = torch.randn(E, D, dtype=torch.float64)
    mq, cq, w_out, b_out = layer._folded(torch.float64, with_edge=True)
    HF = H * F
    assert mq.shape == (HF + H * D, F) and cq.shape == (HF + H * D,) and w_out.shape == (F, HF + 8 + F + H * D)
    ur = x @ mq.T + cq.double()
    u, r = ur[:, :HF].view(N, H, F), ur[:, HF:].view(N, H, D)
    src, dst = ei[0], ei[1]
    logit = (u[dst] * x[src].unsqueeze(1)).sum(-1) + (r[dst] * ea.unsqueeze(1)).sum(-1)        # [E, H]
    alpha = lo.segment_softmax(logit, dst, N)
    z = torch.zeros(N, H, F, dtype=torch.float64).index_add_(0, dst, alpha.unsqueeze(-1) * x[src].unsqueeze(1))
    sw = torch.zeros(N, H, dtype=torch.float64).index_add_(0, dst, alpha)
    m = torch.zeros(N, H, D, dtype=torch.float64).index_add_(0, dst, alpha.unsqueeze(-1) * ea.unsqueeze(1))
    z_aug = torch.cat([z.reshape(N, HF), sw, torch.zeros(N, 8 - H, dtype=torch.float64), x, m.reshape(N, H * D)], 1)
    out = z_aug @ w_out.T + b_out.double()
    p = {k: v.detach() for k, v in layer.state_dict().items()}
    ref = lo.transformer_conv(x, ei, p['lin_query.weight'], p['lin_query.bias'], p['lin_key.weight'], p['lin_key.bias'],
                              p['lin_value.weight'], p['lin_value.bias'], p['lin_skip.weight'], p['lin_skip.bias'],
                              heads=H, edge_attr=ea, we=p['lin_edge.weight'])
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-7)      # _folded hands the bias vectors over in fp32


def test_segment_softmax_and_loops():
    src = torch.tensor([[1.0], [3.0], [2.0], [-1.0]], dtype=torch.float64)
    idx = torch.tensor([0, 0, 2, 2])
    p = lo.segment_softmax(src, idx, 4)
    assert abs(float(p[:2].sum()) - 1) < 1e-12 and abs(float(p[2:].sum()) - 1) < 1e-12
    ei = torch.tensor([[0, 1, 2, 2], [1, 1, 0, 2]])
    assert lo.replace_self_loops(ei, 3).tolist() == [[0, 2, 0, 1, 2], [1, 0, 0, 1, 2]]


def test_flow_gnn_forward_runs_for_all_layer_types():
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    ei = graph(12)
    x = torch.randn(12, 3, dtype=torch.float64)
    for lt in ("GCN", "GAT", "GIN", "Transformer"):
        p = {k: v.double() for k, v in FlowGNN(3, 16, 7, 2, lt).state_dict().items()}
        out = lo.flow_gnn_forward(x, ei, p, lt, training=False)
        assert out.shape == (12, 7) and torch.isfinite(out).all()


@pytest.mark.parametrize("kind", ["GCN", "GAT", "GIN", "Transformer"])
def test_sampled_closure_rows_equal_full_graph_rows(kind):
    """oracle/sampled.py: the one-hop closure sub-problem reproduces the full-graph oracle rows exactly (it is what
    bench.py's parity_check evaluates at cfg3 / cfg4 sizes, where the full fp64 oracle cannot run)."""
    from oracle import sampled
    N, E, F, H = 400, 2400, 16, 4
    rng = np.random.default_rng(5)
    ei = torch.from_numpy(rng.integers(0, N - 3, size=(2, E)))
    ei[1, :40] = ei[0, :40]
    ei[1, 100:180] = 7                                   # hub target
    torch.manual_seed(0)
    x = torch.randn(N, F, dtype=torch.float64)
    g = lambda *s: torch.randn(*s, dtype=torch.float64) * 0.3
    if kind == "GCN":
        p = {"lin.weight": g(F, F), "bias": g(F)}
        full = lo.gcn_conv(x, ei, p["lin.weight"], p["bias"])
    elif kind == "GAT":
        p = {"lin.weight": g(H * F, F), "att_src": g(1, H, F), "att_dst": g(1, H, F), "bias": g(F)}
        full = lo.gat_conv(x, ei, p["lin.weight"], p["att_src"], p["att_dst"], p["bias"], heads=H)
    elif kind == "GIN":
        p = {"nn.0.weight": g(F, F), "nn.0.bias": g(F), "nn.2.weight": g(F, F), "nn.2.bias": g(F), "eps": torch.zeros(1)}
        full = lo.gin_conv(x, ei, lo.gin_mlp(p["nn.0.weight"], p["nn.0.bias"], p["nn.2.weight"], p["nn.2.bias"]))
    else:
        p = {f"lin_{n}.weight": g(H * F, F) for n in ("query", "key", "value")}
        p.update({f"lin_{n}.bias": g(H * F) for n in ("query", "key", "value")})
        p.update({"lin_skip.weight": g(F, F), "lin_skip.bias": g(F)})
        full = lo.transformer_conv(x, ei, p["lin_query.weight"], p["lin_query.bias"], p["lin_key.weight"], p["lin_key.bias"],
                                   p["lin_value.weight"], p["lin_value.bias"], p["lin_skip.weight"], p["lin_skip.bias"], heads=H)
    rows = sampled.pick_rows(N, 60, plane=50, seed=1)
    assert rows.numel() == 60 and int(rows.min()) == 0 and int(rows.max()) == N - 1
    rows = torch.unique(torch.cat([rows, torch.tensor([7, N - 2])]))        # the hub and an isolated node
    nodes, ei_sub, pos = sampled.closure_subgraph(ei, rows, N)
    assert nodes.numel() < N and ei_sub.shape[1] < E
    sub = sampled.layer_rows(kind, p, x[nodes], ei_sub, pos, heads=H)
    torch.testing.assert_close(sub, full[rows], rtol=1e-12, atol=1e-12)
