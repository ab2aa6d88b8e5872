"""K1 parity (bit-exact): device CSR vs the oracle's stable-sort definition."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand_graph(N, E, seed, hub=None, loops=0.1):
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, N, size=(2, E))
    nl = int(E * loops)
    if nl:
        ei[1, :nl] = ei[0, :nl]                      # pre-existing self loops (some duplicated)
    if hub is not None:
        ei[1, E // 2:E // 2 + hub] = 3               # a heavy row (deg > 32 path)
        ei[0, E // 3:E // 3 + hub] = 5               # a heavy source row
    rng.shuffle(ei, axis=1)
    return ei


@pytest.mark.parametrize("N,E,hub", [(50, 400, None), (1000, 6000, 200), (20000, 90000, 5000), (7, 0, None), (1, 3, None)])
@pytest.mark.parametrize("self_loops", [False, True])
@pytest.mark.parametrize("by_source", [False, True])
def test_csr_matches_stable_sort(N, E, hub, self_loops, by_source):
    from gnn_bfs_rans_b200 import ops
    from oracle import builder_oracle as bo
    ei = _rand_graph(N, E, N + E, hub if E else None) if E else np.zeros((2, 0), dtype=np.int64)
    if N == 1:
        ei = np.zeros((2, E), dtype=np.int64)
    rowptr, col, eid, dinv = ops.csr_build(torch.from_numpy(ei).cuda(), N, self_loops, by_source, want_dinv=True)
    eff = bo.effective_edges(ei, N, self_loops)
    r_rowptr, r_col, r_order = bo.csr_by_target(eff, N, by_source)
    np.testing.assert_array_equal(rowptr.cpu().numpy(), r_rowptr)
    np.testing.assert_array_equal(col.cpu().numpy(), r_col)
    # edge ids: kept edges carry their ORIGINAL position, appended loops E+v
    if self_loops:
        keep = np.nonzero(ei[0] != ei[1])[0]
        ids = np.concatenate([keep, E + np.arange(N)])
    else:
        ids = np.arange(E)
    np.testing.assert_array_equal(eid.cpu().numpy(), ids[r_order])
    deg = np.diff(r_rowptr).astype(np.float32)
    ref_dinv = torch.from_numpy(deg).pow(-0.5)
    ref_dinv[ref_dinv == float('inf')] = 0
    assert torch.equal(dinv.cpu(), ref_dinv)          # bit-exact 1/sqrt


def test_perm_and_determinism():
    from gnn_bfs_rans_b200.graph import Graph
    N, E = 5000, 40000
    ei = torch.from_numpy(_rand_graph(N, E, 1, hub=100)).cuda()
    for variant in ("raw", "sl"):
        g = Graph(ei, N)
        a, b, perm = g.csr(variant, False), g.csr(variant, True), g.perm(variant)
        assert torch.equal(a.eid[perm.long()], b.eid)               # same edge at both positions
        # (target of CSR pos) == (col of transposed pos) and vice versa
        rows_a = torch.repeat_interleave(torch.arange(N, device='cuda'), (a.rowptr[1:] - a.rowptr[:-1]).long())
        rows_b = torch.repeat_interleave(torch.arange(N, device='cuda'), (b.rowptr[1:] - b.rowptr[:-1]).long())
        assert torch.equal(rows_a[perm.long()].int(), b.col) and torch.equal(a.col[perm.long()], rows_b.int())
        g2 = Graph(ei, N)
        assert torch.equal(g2.csr(variant, False).col, a.col) and torch.equal(g2.csr(variant, False).eid, a.eid)


def test_out_of_range_edges_raise():
    from gnn_bfs_rans_b200 import ops
    ei = torch.tensor([[0, 1, 9], [1, 0, 2]], device='cuda')
    with pytest.raises(RuntimeError):
        ops.csr_build(ei, 5, False, False, False)


def test_graph_cache_identity():
    from gnn_bfs_rans_b200.graph import graph_of, clear_cache
    clear_cache()
    ei = torch.randint(0, 100, (2, 500), device='cuda')
    g1 = graph_of(ei, 100)
    assert graph_of(ei, 100) is g1
    assert graph_of(ei.clone(), 100) is not g1
    ei[0, 0] = (ei[0, 0] + 1) % 100                   # in-place edit bumps _version -> rebuilt
    assert graph_of(ei, 100) is not g1
