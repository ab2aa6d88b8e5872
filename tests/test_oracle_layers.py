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
