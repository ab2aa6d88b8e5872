"""Multi-GPU message passing (the reference has no counterpart — it is single-process).

One process per GPU.  Cells are partitioned by recursive coordinate bisection (RCB) of the cell
centres; each rank owns its cells plus a 1-ring halo (ghost cells = sources of edges whose target it
owns).  Local node ids are [owned | ghosts grouped by owning rank, ascending global id]; the local
CSR only has rows for owned targets.  One exchange step per layer: the rows a peer needs are packed
by a libb2g.so gather kernel and shipped with NCCL send/recv over NVLink; the backward is the
reverse exchange with a scatter-add into the owners.  Global reductions (BatchNorm statistics, loss
means, the flat gradient all-reduce) keep single-GPU parity.

The plan (who sends which rows to whom) is computed WITHOUT communication: every rank derives both
sides from the same global partition vector, so send and receive orders agree by construction."""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------ RCB
def rcb_partition(centers: torch.Tensor, parts: int) -> torch.Tensor:
    """Recursive coordinate bisection.  centers [N,3] -> part id int64 [N] in [0, parts); parts = 2^k.
    Each split cuts the longest extent of the current box at the median (ties by cell id, so the two
    halves differ by at most one cell)."""
    if parts & (parts - 1):
        raise ValueError("rcb_partition: parts must be a power of two")
    N = centers.shape[0]
    part = torch.zeros(N, dtype=torch.int64, device=centers.device)
    groups = [torch.arange(N, device=centers.device)]
    p = 1
    while p < parts:
        nxt = []
        for gi, idx in enumerate(groups):
            c = centers[idx]
            ext = c.max(0).values - c.min(0).values if idx.numel() else torch.zeros(3, device=centers.device)
            ax = int(torch.argmax(ext))
            order = torch.argsort(c[:, ax], stable=True)
            half = (idx.numel() + 1) // 2
            lo, hi = idx[order[:half]], idx[order[half:]]
            lo, hi = lo.sort().values, hi.sort().values
            nxt += [lo, hi]
        groups = nxt
        p *= 2
    for gi, idx in enumerate(groups):
        part[idx] = gi
    return part


# ------------------------------------------------------------------------------------------ plan
class Partition:
    """One rank's share of a partitioned graph + its halo exchange plan."""

    def __init__(self, rank: int, world: int, n_owned: int, owned_global: Optional[torch.Tensor],
                 edge_index: torch.Tensor, send_idx: List[torch.Tensor], recv_counts: List[int],
                 ghost_global: Optional[torch.Tensor], n_edges_raw_owned: int, n_loops_owned: int):
        self.rank, self.world = rank, world
        self.n_owned = int(n_owned)
        self.recv_counts = [int(c) for c in recv_counts]
        self.n_ghost = sum(self.recv_counts)
        self.n_local = self.n_owned + self.n_ghost
        self.owned_global, self.ghost_global = owned_global, ghost_global
        self.edge_index = edge_index                      # local ids; every target < n_owned
        self.send_idx = send_idx                          # per peer: int32 local (owned) rows to send
        self.send_counts = [int(t.numel()) for t in send_idx]
        self._send_all = torch.cat(send_idx) if any(self.send_counts) else None
        self._e_raw, self._loops = int(n_edges_raw_owned), int(n_loops_owned)
        self._dinv_ready = False

    # edges a layer aggregates for the OWNED targets (the unit of the throughput metric)
    def aggregated_edges(self, layer_type: str) -> int:
        if layer_type in ("GCN", "GAT"):
            return self._e_raw - self._loops + self.n_owned       # self loops replaced: E_sl
        return self._e_raw

    # ---------------------------------------------------------------- exchange
    def exchange(self, x_full: torch.Tensor, gather: Optional[Callable] = None) -> torch.Tensor:
        """Fill the ghost rows x_full[n_owned:] with the owners' rows (in place).  `gather(x, idx)` packs
        rows; the default is the libb2g.so kernel (CUDA only)."""
        if self.world == 1 or (self.n_ghost == 0 and not any(self.send_counts)):
            return x_full
        if gather is None:
            from . import ops
            gather = ops.rows_gather
        pack = gather(x_full, self._send_all) if self._send_all is not None else x_full[:0]
        reqs, off_s, off_r = [], 0, self.n_owned
        ops_list = []
        for peer in range(self.world):
            ns, nr = self.send_counts[peer], self.recv_counts[peer]
            if ns:
                ops_list.append(dist.P2POp(dist.isend, pack[off_s:off_s + ns], peer))
            if nr:
                ops_list.append(dist.P2POp(dist.irecv, x_full[off_r:off_r + nr], peer))
            off_s += ns
            off_r += nr
        if ops_list:
            reqs = dist.batch_isend_irecv(ops_list)
            for r in reqs:
                r.wait()
        return x_full

    def exchange_reverse_add(self, g_full: torch.Tensor, scatter_add: Optional[Callable] = None) -> torch.Tensor:
        """Backward of `exchange`: ghost-row gradients travel back and are added into the owners' rows."""
        if self.world == 1 or (self.n_ghost == 0 and not any(self.send_counts)):
            return g_full
        if scatter_add is None:
            from . import ops
            scatter_add = ops.rows_scatter_add
        n_send = sum(self.send_counts)
        back = g_full.new_empty((n_send, g_full.shape[1]))
        ops_list, off_s, off_r = [], 0, self.n_owned
        for peer in range(self.world):
            ns, nr = self.send_counts[peer], self.recv_counts[peer]
            if nr:
                ops_list.append(dist.P2POp(dist.isend, g_full[off_r:off_r + nr].contiguous(), peer))
            if ns:
                ops_list.append(dist.P2POp(dist.irecv, back[off_s:off_s + ns], peer))
            off_s += ns
            off_r += nr
        if ops_list:
            for r in dist.batch_isend_irecv(ops_list):
                r.wait()
        # a row may be needed by several peers: add peer by peer so each call sees unique indices
        off = 0
        for peer in range(self.world):
            ns = self.send_counts[peer]
            if ns:
                scatter_add(g_full, self.send_idx[peer], back[off:off + ns])
            off += ns
        return g_full

    # ---------------------------------------------------------------- graph glue
    def prepare_graph(self, exchange: Optional[Callable] = None):
        """Build the cached Graph of the local edge list and patch the GCN deg^-1/2 of the ghost rows
        with their owners' values (a ghost's local row has no incoming edges).  `exchange(buf)` fills the
        ghost rows of a [n_local, 4] fp32 buffer in place (default: self.exchange over NCCL)."""
        from .graph import graph_of
        g = graph_of(self.edge_index, self.n_local)
        # the "ghost deg^-1/2 patched" flag lives on the Graph object: if the cache ever hands out a fresh Graph for this
        # edge_index (eviction), it is patched again instead of being trusted
        if self.world > 1 and not (self._dinv_ready and getattr(g, "_ghost_dinv_patched", False)):
            dinv = g.dinv()
            buf = torch.zeros((self.n_local, 4), dtype=torch.float32, device=dinv.device)
            buf[:, 0] = dinv
            (exchange or self.exchange)(buf)
            dinv[self.n_owned:] = buf[self.n_owned:, 0]
            g._ghost_dinv_patched = True
            g._pinned = True                 # graph.graph_of never evicts it while edge_index lives
        self._dinv_ready = True
        self._graph = g          # strong reference: the patched Graph lives as long as the partition
        return g

    def wrap_forward(self, layer):
        """fn(x_owned_or_full, edge_index) -> out for the owned rows, with the halo exchange in front.
        Ghost projections are recomputed locally (exchange x, width F) rather than shipping H*C-wide rows."""
        self.prepare_graph()
        if self.world == 1:
            return lambda x, ei: layer(x, ei)
        holder = {}
        from .nn import GCNConv
        # GCNConv inference: the exchange (pack + NCCL send/recv on a side stream) runs behind the projection of the owned
        # rows; bit-equal with the blocking path (2-GPU test).  B2G_HALO_OVERLAP=0 selects the blocking exchange for A/B runs.
        overlapped = (self._gcn_forward_overlapped(layer, holder)
                      if isinstance(layer, GCNConv) and os.environ.get("B2G_HALO_OVERLAP", "1") != "0" else None)

        def fwd(x, ei):
            if overlapped is not None and x.is_cuda and not (torch.is_grad_enabled() and (
                    x.requires_grad or any(p.requires_grad for p in layer.parameters()))):
                return overlapped(x, ei)
            if x.shape[0] == self.n_local:
                xf = x
            else:
                xf = holder.get("buf")
                if xf is None or xf.shape[1] != x.shape[1] or xf.dtype != x.dtype:
                    xf = holder["buf"] = x.new_empty((self.n_local, x.shape[1]))
                xf[:self.n_owned] = x
            self.exchange(xf)
            return layer(xf, ei)[:self.n_owned]

        return fwd


    def _gcn_forward_overlapped(self, layer, holder):
        """Inference GCNConv on this rank's share with the halo exchange hidden behind the projection of the owned rows:
            side stream : pack + NCCL send/recv of the boundary rows of x          (9.6 MB per rank per layer at cfg4)
            main stream : xs[owned] = dinv * (x[owned] W^T)                        (the K6 GEMM, 1.6 ms at cfg4)
            then        : xs[ghosts] = dinv_ghost * (x[ghosts] W^T)  (a few thousand rows), aggregation over the local CSR.
        Same kernels and bits as `layer(x_full, edge_index)[:n_owned]` after a blocking exchange."""
        from . import ops
        part = self

        def ghost_ranges(g, n0):
            """Row ranges [0, a) and [b, n0) that contain every owned row reading a ghost row (cached on the graph); None when
            they are not two thin end ranges (general partitions: the exchange is then waited for before the kernel)."""
            hit = getattr(g, "_ghost_ranges", False)
            if hit is not False:
                return hit
            csr = g.csr("sl", False)
            pos = torch.nonzero(csr.col[:int(csr.rowptr[n0])] >= n0).squeeze(1)
            res = None
            if pos.numel():
                rows = torch.unique(torch.searchsorted(csr.rowptr[:n0 + 1].long(), pos, right=True) - 1)
                mid = n0 // 2
                lo, hi = rows[rows < mid], rows[rows >= mid]
                a = int(lo.max()) + 1 if lo.numel() else 0
                b = int(hi.min()) if hi.numel() else n0
                if a + (n0 - b) <= n0 // 4:
                    res = (a, b)
            else:
                res = (0, n0)
            g._ghost_ranges = res
            return res

        def fwd(x, ei):
            n0, nl = part.n_owned, part.n_local
            g = part._graph
            csr = g.csr("sl", False)
            dinv = g.dinv()
            if x.shape[0] == nl:
                xf = x
            else:
                xf = holder.get("buf")
                if xf is None or xf.shape[1] != x.shape[1] or xf.dtype != x.dtype:
                    xf = holder["buf"] = x.new_empty((nl, x.shape[1]))
                xf[:n0] = x
            side = holder.get("side")
            if side is None:
                side = holder["side"] = torch.cuda.Stream(x.device)
            cur = torch.cuda.current_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                part.exchange(xf)
                done = torch.cuda.Event()
                done.record(side)
            w = layer.lin.weight if layer.lin.weight.dtype == x.dtype else layer.lin.weight.to(x.dtype)
            if (os.environ.get("B2G_GCN_PATH", "") == "fused"
                    and ops.segw_gemm_supported(nl, x.shape[1], layer.out_channels, x.dtype)):
                # fused aggregation + projection (csrc/gcn_fused.cu) on the raw rows: the interior rows (no ghost sources) run
                # while the boundary rows of x travel; the two thin end ranges that read ghosts follow the exchange
                out = x.new_empty((n0, layer.out_channels))
                rng = ghost_ranges(g, n0)

                def run(r0, r1):
                    if r1 > r0:
                        ops.segw_gemm(xf, csr.rowptr[r0:r1 + 1], csr.col, r1 - r0, w, layer.bias, col_scale=dinv,
                                      row_scale=dinv[r0:r1], band=g.band(), out=out[r0:r1])
                if rng is None:
                    cur.wait_event(done)
                    run(0, n0)
                else:
                    run(rng[0], rng[1])
                    cur.wait_event(done)
                    run(0, rng[0])
                    run(rng[1], n0)
                return out
            xs = x.new_empty((nl, layer.out_channels))
            # B2G_HALO_RESERVE_SMS=r: the persistent GEMM leaves r SMs to the pack kernel and NCCL's send / recv CTAs on the side
            # stream (its CTAs take a whole SM's shared memory).  Measured at N = 2: 4.427 (r = 0) / 4.441 (8) / 4.468 ms (16) — the
            # 25 MB exchange is too short for the reserve to pay, so the default is 0
            ops.linear_fwd(xf[:n0], w, None, row_scale=dinv[:n0], out=xs[:n0],
                           reserve_sms=int(os.environ.get("B2G_HALO_RESERVE_SMS", "0")) if part.world > 1 else 0)
            cur.wait_event(done)
            if nl > n0:
                ops.linear_fwd(xf[n0:], w, None, row_scale=dinv[n0:], out=xs[n0:])
            bias = layer.bias.float() if layer.bias is not None else None
            return ops.seg_sum(xs, csr.rowptr, csr.col, n0, dinv, None, 0.0, None, bias, band=g.band())

        return fwd


class HaloFn(torch.autograd.Function):
    """Differentiable halo exchange: x_owned [n_owned,F] -> x_full [n_local,F]."""

    @staticmethod
    def forward(ctx, x_owned, part: Partition):
        xf = x_owned.new_empty((part.n_local, x_owned.shape[1]))
        xf[:part.n_owned] = x_owned
        part.exchange(xf)
        ctx.part = part
        return xf

    @staticmethod
    def backward(ctx, g_full):
        part = ctx.part
        g = g_full.contiguous().clone()
        part.exchange_reverse_add(g)
        return g[:part.n_owned], None


def build_partition(edge_index: torch.Tensor, part: torch.Tensor, rank: int, world: int,
                    device=None) -> Partition:
    """General (unstructured) case: every rank holds the global edge_index [2,E] and the partition
    vector [N] (e.g. from rcb_partition) and cuts out its share.  Local edge order = global edge order."""
    src, dst = edge_index[0], edge_index[1]
    N = part.numel()
    mine = part[dst] == rank
    src_m, dst_m = src[mine], dst[mine]
    owned = torch.nonzero(part == rank).squeeze(1)                       # ascending global ids
    n_owned = owned.numel()
    ghost_mask = part[src_m] != rank
    ghosts = torch.unique(src_m[ghost_mask])                              # ascending
    gp = part[ghosts]
    order = torch.argsort(gp, stable=True)                                # group by owner, keep id order
    ghosts = ghosts[order]
    recv_counts = torch.bincount(part[ghosts], minlength=world).tolist() if ghosts.numel() else [0] * world
    g2l = torch.full((N,), -1, dtype=torch.int64, device=edge_index.device)
    g2l[owned] = torch.arange(n_owned, device=edge_index.device)
    g2l[ghosts] = n_owned + torch.arange(ghosts.numel(), device=edge_index.device)
    ei_local = torch.stack([g2l[src_m], g2l[dst_m]])
    # what I must send to peer q: my owned cells that are sources of edges whose target q owns
    send_idx = []
    for q in range(world):
        if q == rank:
            send_idx.append(torch.zeros(0, dtype=torch.int32, device=edge_index.device))
            continue
        sel = (part[dst] == q) & (part[src] == rank)
        need = torch.unique(src[sel])                                     # ascending global == q's recv order
        send_idx.append(g2l[need].to(torch.int32))
    n_loops = int((src_m == dst_m).sum())
    dev = device if device is not None else edge_index.device
    return Partition(rank, world, n_owned, owned.to(dev), ei_local.contiguous().to(dev),
                     [s.to(dev) for s in send_idx], recv_counts, ghosts.to(dev), int(src_m.numel()), n_loops)


def slab_partition_hex(nx: int, ny: int, nz: int, world: int, rank: int, device) -> Partition:
    """Weak-scaling bench mesh: a hex block nx x ny x (nz*world), which RCB cuts into `world` slabs of
    nz planes along z.  Each rank generates only its own slab (OpenFOAM-style faces -> device builder
    mode A) plus the edges arriving from the neighbouring slabs' boundary planes."""
    from . import ops
    from .synthetic import hex_mesh_faces
    dev = torch.device(device)
    owner, nei = hex_mesh_faces(nx, ny, nz, device=dev)
    N = nx * ny * nz
    ei = ops.build_graph_edges(owner, nei, 1, None, N, N)                 # [2, E_block], bit-exact builder
    del owner, nei
    plane = nx * ny
    pid = torch.arange(plane, dtype=torch.int64, device=dev)
    send_idx = [torch.zeros(0, dtype=torch.int32, device=dev) for _ in range(world)]
    recv_counts = [0] * world
    extra, goff = [], N
    if rank > 0:                                                           # ghosts from the slab below
        extra.append(torch.stack([goff + pid, pid]))                       # ghost(below c) -> c, z_local = 0
        send_idx[rank - 1] = pid.to(torch.int32)                           # my bottom plane goes down
        recv_counts[rank - 1] = plane
        goff += plane
    if rank < world - 1:                                                   # ghosts from the slab above
        top = (nz - 1) * plane + pid
        extra.append(torch.stack([goff + pid, top]))
        send_idx[rank + 1] = top.to(torch.int32)
        recv_counts[rank + 1] = plane
        goff += plane
    n_raw = ei.shape[1] + sum(e.shape[1] for e in extra)
    if extra:
        ei = torch.cat([ei] + extra, dim=1).contiguous()
    return Partition(rank, world, N, None, ei, send_idx, recv_counts, None, n_raw, 0)


# ------------------------------------------------------------------------------------------ training glue
def allreduce_gradients(params, world: int):
    """One flat all-reduce (SUM) over every parameter gradient (~0.4-2 M parameters -> a
    single bucket), then scatter back.  Loss terms must already be normalised by GLOBAL counts."""
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def flow_forward_partitioned(model, x_owned: torch.Tensor, part: Partition, group=True,
                             checkpoint_layers: bool = False) -> torch.Tensor:
    """`FlowGNN.forward` (gnn_model.py:159-197) for ONE rank's share of a partitioned mesh: per layer a differentiable
    halo exchange of the layer input (HaloFn), the layer on the local graph, and residual + BatchNorm + ReLU + dropout
    on the owned rows with the BatchNorm statistics combined over all ranks (csrc/bn.cu + all_gather / all_reduce of
    [C]-sized vectors).  Output rows = the owned cells.  With the loss normalised by GLOBAL counts and
    `allreduce_gradients` afterwards, a step equals the single-process step on the whole mesh (tested on 2 GPUs).

    checkpoint_layers: keep only each block's input and re-run the block (exchange included: every rank does, in the
    same order) in backward — what lets 12.5 M cells per GPU (cfg5: 100 M cells on 8 GPUs) fit in 180 GB.  The torch RNG
    state is restored for the re-run, so the attention / glue dropout masks are the same; BatchNorm running statistics
    are updated by the first run only."""
    from . import functional as Fn
    from . import ops
    part.prepare_graph()
    h = model.input_proj(x_owned)
    if model.use_batch_norm and not ops.bn_supported(h):
        raise RuntimeError("b2g: partitioned BatchNorm needs a CUDA [N, C] input with 16-byte-multiple rows")
    runs = {}

    def block(h_in, i):
        layer = model.gnn_layers[i]
        hf = HaloFn.apply(h_in, part)
        h_new = layer(hf, part.edge_index)[:part.n_owned]
        if not model.use_batch_norm:
            return model.dropout(torch.relu(h_in + h_new))
        bn = model.batch_norms[i].module
        rerun = runs.get(i, 0) > 0                       # second execution of this block = the checkpoint re-run
        runs[i] = runs.get(i, 0) + 1
        saved = (bn.momentum, bn.num_batches_tracked.clone() if bn.num_batches_tracked is not None else None)
        if rerun:
            bn.momentum = 0.0                            # running statistics were updated by the first run
        try:
            return Fn.batch_norm(h_in, h_new, bn, relu=True, p_drop=model.dropout.p,
                                 group=group if part.world > 1 else None)
        finally:
            if rerun:
                bn.momentum = saved[0]
                if saved[1] is not None:
                    bn.num_batches_tracked.copy_(saved[1])

    for i in range(len(model.gnn_layers)):
        if checkpoint_layers and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint
            h = checkpoint(block, h, i, use_reentrant=False)
        else:
            h = block(h, i)
    return model.output_proj(h)
