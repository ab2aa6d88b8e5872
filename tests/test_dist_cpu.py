"""N>1 host-side logic on CPU (gloo, world_size 2 and 4 via virtual ranks): RCB partitioner, halo plan,
exchange and its reverse.  The exchange is driven with plain torch indexing supplied BY THE TEST
(`gather=` / `scatter_add=` hooks) — the product's default is the libb2g.so kernel and needs CUDA."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_bfs_rans_b200.distributed import Partition, build_partition, rcb_partition
from oracle import layers_oracle as lo


def grid_edges(nx, ny, nz):
    ids = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    pairs = []
    for a, b in ((ids[:, :, :-1], ids[:, :, 1:]), (ids[:, :-1, :], ids[:, 1:, :]), (ids[:-1], ids[1:])):
        pairs.append(np.stack([a.ravel(), b.ravel()]))
    p = np.concatenate(pairs, axis=1)
    ei = np.concatenate([p, p[::-1]], axis=1)
    ei = np.concatenate([ei, np.array([[3, 3, 7], [3, 3, 7]])], axis=1)     # pre-existing loops, duplicated
    return torch.from_numpy(ei).long()


def centers(nx, ny, nz):
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return torch.from_numpy(np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float64))


def test_rcb_balanced_and_spatial():
    c = centers(8, 6, 4)
    for P in (1, 2, 4, 8):
        part = rcb_partition(c, P)
        cnt = torch.bincount(part, minlength=P)
        assert int(cnt.max() - cnt.min()) <= 1 and int(cnt.sum()) == 192
    p2 = rcb_partition(c, 2)
    assert bool((c[p2 == 0][:, 0].max() < c[p2 == 1][:, 0].min()))          # first cut: longest extent (x)
    with pytest.raises(ValueError):
        rcb_partition(c, 3)


def test_partition_plan_consistency_virtual_ranks():
    """All ranks' plans computed in one process: send lists of p to q == ghosts q expects from p."""
    ei, c = grid_edges(6, 5, 4), centers(6, 5, 4)
    for P in (2, 4):
        part = rcb_partition(c, P)
        plans = [build_partition(ei, part, r, P) for r in range(P)]
        assert sum(p.n_owned for p in plans) == 120
        assert sum(p.edge_index.shape[1] for p in plans) == ei.shape[1]      # every edge has one owner
        for q in range(P):
            off = plans[q].n_owned
            for p in range(P):
                n = plans[q].recv_counts[p]
                want = plans[q].ghost_global[off - plans[q].n_owned: off - plans[q].n_owned + n]
                sent = plans[p].owned_global[plans[p].send_idx[q].long()]
                assert torch.equal(want, sent), (P, p, q)
                off += n
            assert bool((plans[q].edge_index[1] < plans[q].n_owned).all())   # targets are owned
        # partitioned GCN aggregation == monolithic (virtual exchange by index copies)
        x = torch.randn(120, 8, dtype=torch.float64)
        W, b = torch.randn(8, 8, dtype=torch.float64), torch.randn(8, dtype=torch.float64)
        ref = lo.gcn_conv(x, ei, W, b)
        ei_sl, w = lo.gcn_norm(ei, 120, torch.float64)
        for q in range(P):
            pl = plans[q]
            gl = torch.cat([pl.owned_global, pl.ghost_global])
            xl = x[gl]
            # local normalisation needs the GLOBAL degree of ghosts: take dinv from the monolithic graph
            deg = torch.zeros(120, dtype=torch.float64).scatter_add_(0, ei_sl[1], torch.ones(ei_sl.shape[1], dtype=torch.float64))
            dinv = deg.pow(-0.5)[gl]
            nl = pl.edge_index[:, pl.edge_index[0] != pl.edge_index[1]]
            h = xl @ W.T
            out = torch.zeros(pl.n_owned, 8, dtype=torch.float64)
            out.index_add_(0, nl[1], (dinv[nl[0]] * dinv[nl[1]]).unsqueeze(1) * h[nl[0]])
            out += (dinv[:pl.n_owned] ** 2).unsqueeze(1) * h[:pl.n_owned] + b
            torch.testing.assert_close(out, ref[pl.owned_global], rtol=1e-12, atol=1e-12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei, c = grid_edges(6, 5, 4), centers(6, 5, 4)
        part = rcb_partition(c, world)
        pl = build_partition(ei, part, rank, world)
        torch.manual_seed(0)
        x = torch.randn(120, 8)
        xf = torch.zeros(pl.n_local, 8)
        xf[:pl.n_owned] = x[pl.owned_global]
        pl.exchange(xf, gather=lambda t, idx: t[idx.long()].contiguous())
        ok_fwd = torch.equal(xf[pl.n_owned:], x[pl.ghost_global])
        # reverse: every ghost row carries 1.0 -> owners receive (number of ranks that ghost the row)
        g = torch.zeros(pl.n_local, 8)
        g[pl.n_owned:] = 1.0

        def sadd(t, idx, src):
            t.index_add_(0, idx.long(), src)
        pl.exchange_reverse_add(g, scatter_add=sadd)
        cnt = torch.zeros(120)
        for r in range(world):
            other = build_partition(ei, part, r, world)
            cnt[other.ghost_global] += 1
        ok_bwd = torch.equal(g[:pl.n_owned, 0], cnt[pl.owned_global])
        # flat gradient all-reduce
        from gnn_bfs_rans_b200.distributed import allreduce_gradients
        w = torch.nn.Parameter(torch.ones(3))
        w.grad = torch.full((3,), float(rank + 1))
        allreduce_gradients([w], world)
        ok_ar = torch.equal(w.grad, torch.full((3,), float(sum(range(1, world + 1)))))
        q.put((rank, ok_fwd, ok_bwd, ok_ar))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] and r[3] for r in res), res


def _stats_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnn_bfs_rans_b200.functional import combine_batch_stats
        torch.manual_seed(0)
        full = torch.randn(1000, 16, dtype=torch.float64) * 3 + 5
        cuts = [0, 130, 1000] if world == 2 else [0, 100, 400, 401, 1000]     # very uneven shares, one of a single row
        mine = full[cuts[rank]:cuts[rank + 1]]
        stats, n = combine_batch_stats(mine.mean(0), mine.var(0, unbiased=False), mine.shape[0], 1e-5)
        ref_mean, ref_var = full.mean(0), full.var(0, unbiased=False)
        err = max(float((stats[0].double() - ref_mean).abs().max() / ref_mean.abs().max()),
                  float((stats[2].double() - ref_var).abs().max() / ref_var.abs().max()),
                  float((stats[1].double() - (ref_var + 1e-5).rsqrt()).abs().max() / (ref_var + 1e-5).rsqrt().abs().max()))
        q.put((rank, n, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_batchnorm_statistics_combine_to_the_global_ones(world):
    """The synchronised BatchNorm of flow_forward_partitioned: per-rank (n, mean, var) of uneven shares combine to the
    statistics of the whole batch on every rank (gloo, CPU)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, n, err in res:
        assert n == 1000 and err < 1e-6, res
