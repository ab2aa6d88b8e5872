"""Layer forward from HOST buffers with the PCIe copies overlapped with the kernels.

The reference moves a whole batch with `batch.to(device)` (/root/reference/train.py:167), runs the layer
(`gnn_model.py:166`) and reads the result back (`inference.py:87` `.cpu()`): host->device copy, compute and
device->host copy strictly one after the other.  For the 10 M-cell case that is 6.1 GB in and 5.1 GB out over a
~55 GB/s link: ~200 ms of copies around 4 ms of kernels.  PCIe is full duplex and the layer is row-local up to the
index band of the mesh, so the three phases pipeline:

    copy stream in : edge_index, then x in row chunks              (host -> device)
    compute stream : CSR build; per chunk the K6 Linear (+ fused dinv) as soon as the chunk has landed, and the
                     K2 aggregation of every row chunk whose neighbour rows [r0 - band, r1 + band) are projected
    copy stream out: each aggregated chunk                          (device -> host)

Same kernels, same arithmetic and the same result bits as `layer(x.cuda(), edge_index.cuda()).cpu()`: the chunking
only changes WHEN a row is computed (each row is computed exactly once, in the same per-row summation order)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .graph import Graph
from .nn import GATConv, GCNConv, GINConv, TransformerConv, _cached_fold

_STREAMS = {}


def _streams(dev):
    s = _STREAMS.get(dev)
    if s is None:
        s = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        _STREAMS[dev] = s
    return s


def bind_host_to_gpu_numa(device=None) -> bool:
    """Pin the calling process to the CPUs that are local to `device`'s PCIe root (sysfs `local_cpulist`), so that the
    pinned host buffers it allocates afterwards are first-touched on the GPU's NUMA node.  With one process per GPU on a
    multi-socket box the host<->device copies otherwise cross the socket interconnect (measured at N = 8: 8 ranks moved
    90 GB per step at 142 GB/s in total).  Returns False (and changes nothing) when the topology cannot be read."""
    import os
    try:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        p = torch.cuda.get_device_properties(dev)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return False
        os.sched_setaffinity(0, ids)
        return True
    except Exception:
        return False


def _pinned(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.is_cuda:
        raise RuntimeError(f"b2g.streaming: {what} must be a host tensor")
    return t if t.is_pinned() else t.pin_memory()


@torch.no_grad()
def gcn_forward_host(layer: GCNConv, x_host: torch.Tensor, edge_index_host: torch.Tensor,
                     out_host: Optional[torch.Tensor] = None, rows_per_chunk: int = 1 << 18,
                     partition=None) -> torch.Tensor:
    """GCNConv.forward(x, edge_index) (gnn_model.py:63,166) for host-resident `x` [N,F] and `edge_index` [2,E]
    (int64); returns the host tensor [N, out_channels] (pinned).  Inference only (no autograd graph).

    partition (distributed.Partition, optional): `x_host` / the result hold this rank's OWNED rows and `edge_index_host`
    is the rank's local edge list (ghost sources numbered after the owned rows).  The projected ghost rows are fetched
    from their owners with ONE halo exchange after the last owned chunk is projected; only the row chunks that read a
    ghost wait for it."""
    if not isinstance(layer, GCNConv):
        raise NotImplementedError("b2g.streaming: pipelined host forward exists for GCNConv; use layer(x.cuda(), ei.cuda()) otherwise")
    w = layer.lin.weight
    if not w.is_cuda:
        raise RuntimeError("b2g.streaming: the layer must live on a CUDA device (no CPU fallback)")
    dev = w.device
    if x_host.dim() != 2 or edge_index_host.dim() != 2 or edge_index_host.shape[0] != 2:
        raise ValueError("x must be [N,F] and edge_index [2,E]")
    if x_host.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"b2g: x dtype {x_host.dtype} not supported (float32 / bfloat16)")
    hx = _pinned(x_host.contiguous(), "x")
    hei = _pinned(edge_index_host.long().contiguous(), "edge_index")
    N, F_in = hx.shape
    n_local = N
    multi = partition is not None and partition.world > 1
    if partition is not None:
        if N != partition.n_owned:
            raise ValueError(f"x_host must hold the partition's {partition.n_owned} owned rows, got {N}")
        n_local = partition.n_local
    F_out = layer.out_channels
    if out_host is None:
        out_host = torch.empty((N, F_out), dtype=hx.dtype).pin_memory()
    elif out_host.is_cuda or not out_host.is_pinned() or out_host.shape != (N, F_out) or out_host.dtype != hx.dtype:
        raise ValueError("out_host must be a pinned host tensor [N, out_channels] of x's dtype")
    if N == 0:
        return out_host
    s_in, s_cmp, s_out = _streams(dev)
    cur = torch.cuda.current_stream(dev)
    for s in (s_in, s_cmp, s_out):
        s.wait_stream(cur)
    R = max(1024, int(rows_per_chunk))
    bounds = [(r0, min(r0 + R, N)) for r0 in range(0, N, R)]

    # ---- copy stream in: edge_index first (the CSR build needs all of it), then x chunk by chunk
    with torch.cuda.stream(s_in):
        dei = torch.empty(hei.shape, dtype=torch.int64, device=dev)
        dei.copy_(hei, non_blocking=True)
        ev_ei = torch.cuda.Event()
        ev_ei.record(s_in)
        dx = torch.empty((n_local, F_in), dtype=hx.dtype, device=dev)      # [owned | ghost rows (fused path: filled by the exchange)]
        ev_x = []
        for r0, r1 in bounds:
            dx[r0:r1].copy_(hx[r0:r1], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_in)
            ev_x.append(e)

    with torch.cuda.stream(s_cmp):
        for t in (dei, dx):
            t.record_stream(s_cmp)
        s_cmp.wait_event(ev_ei)
        g = Graph(dei, n_local)
        csr = g.csr("sl", False)
        dinv = g.dinv()
        ghost_chunks = set()
        if multi:                              # ghost rows have no in-edges locally: their deg^-1/2 comes from the owners
            buf = torch.zeros((n_local, 4), dtype=torch.float32, device=dev)
            buf[:, 0] = dinv
            partition.exchange(buf)
            dinv[N:] = buf[N:, 0]
            pos = torch.nonzero(csr.col >= N).squeeze(1)          # CSR positions that read a ghost row
            if pos.numel():
                rows = torch.searchsorted(csr.rowptr.long(), pos, right=True) - 1
                ghost_chunks = set(torch.unique(rows // R).tolist())
        # rows [r0, r1) read owned rows within +-band of themselves (ghost reads are tracked per chunk above):
        # one device reduction + host sync; the x copies are already queued
        dsrc, ddst = dei[0], dei[1]
        band = int(torch.where(dsrc < N, (dsrc - ddst).abs(), torch.zeros_like(dsrc)).max()) if dei.shape[1] else 0
        wt = w if w.dtype == hx.dtype else w.to(hx.dtype)
        bias = layer.bias.float() if layer.bias is not None else None
        import os
        # opt-in B2G_GCN_PATH=fused (bf16, F = 256; csrc/gcn_fused.cu): every chunk is ONE kernel on the raw rows of x (aggregation + projection), the
        # same arithmetic as GCNConv.forward on the whole graph; otherwise project each chunk as it lands, aggregate when ready
        fused = (os.environ.get("B2G_GCN_PATH", "") == "fused" and ops.segw_gemm_supported(n_local, F_in, F_out, hx.dtype))
        xs = dx if fused else torch.empty((n_local, F_out), dtype=hx.dtype, device=dev)
        out = torch.empty((N, F_out), dtype=hx.dtype, device=dev)
        out.record_stream(s_out)
        done_rows, pending = 0, list(range(len(bounds)))           # rows projected so far / chunks not yet aggregated

        def aggregate(c):
            r0, r1 = bounds[c]
            if fused:
                ops.segw_gemm(xs, csr.rowptr[r0:r1 + 1], csr.col, r1 - r0, wt, bias, col_scale=dinv, row_scale=dinv[r0:r1],
                              out=out[r0:r1])
            else:
                ops.seg_sum(xs, csr.rowptr[r0:r1 + 1], csr.col, r1 - r0, dinv[r0:r1], None, 0.0, None, bias,
                            out=out[r0:r1], band=g.band())
            e = torch.cuda.Event()
            e.record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                out_host[r0:r1].copy_(out[r0:r1], non_blocking=True)

        def aggregate_ready(ghosts_in: bool):
            for c in list(pending):
                r0, r1 = bounds[c]
                if min(r1 + band, N) > done_rows or (c in ghost_chunks and not ghosts_in):
                    continue
                aggregate(c)
                pending.remove(c)

        for c, (r0, r1) in enumerate(bounds):
            s_cmp.wait_event(ev_x[c])
            if not fused:
                ops.linear_fwd(dx[r0:r1], wt, None, row_scale=dinv[r0:r1], out=xs[r0:r1])
            done_rows = r1
            aggregate_ready(False)
        if multi:
            partition.exchange(xs)             # boundary rows (projected + dinv-scaled, or raw when fused) -> the neighbours' ghost rows
        aggregate_ready(True)
    cur.wait_stream(s_out)
    cur.wait_stream(s_cmp)
    s_out.synchronize()                        # the result is host memory: return only when it is complete
    return out_host


# ------------------------------------------------------------------------------------------ all four layer types
class _Stages:
    """What a layer type does per row chunk of the host pipeline: `project(r0, r1)` as soon as the chunk of x has landed,
    `finish(r0, r1) -> rows [r0, r1) of the output` once every neighbour row of the chunk is projected.  Same kernels and
    arithmetic as the layer's forward on the whole graph; the CSR positions, the attention weights and every gathered
    matrix are indexed globally, only the target rows are a range."""
    variant = "sl"

    def __init__(self, layer, dx, g: Graph):
        self.layer, self.dx, self.g = layer, dx, g
        self.csr = g.csr(self.variant, False)
        self.N = dx.shape[0]

    def project(self, r0, r1):
        pass


class _GINStages(_Stages):
    variant = "raw"

    def finish(self, r0, r1):
        L = self.layer
        eps = float(L.eps.item()) if not hasattr(self, "_eps") else self._eps
        self._eps = eps
        h = ops.seg_sum(self.dx, self.csr.rowptr[r0:r1 + 1], self.csr.col, r1 - r0, None, None, 1.0 + eps, self.dx[r0:r1], None)
        return L._mlp(h)


class _GATStages(_Stages):
    def __init__(self, layer, dx, g):
        super().__init__(layer, dx, g)
        L = layer
        H, C, F = L.heads, L.out_channels, L.in_channels
        if L.concat or H != 4 or not ops.gatw_gemm_supported(self.N, H, F, C, dx.dtype):
            raise NotImplementedError
        ps = (L.lin.weight, L.att_src, L.att_dst)
        wc, self.v = _cached_fold(L, 'wc_v', ps, dx.dtype, lambda: L._wc_v(dx.dtype))
        self.wp = wc.view(C, H, F // 64, 64).permute(0, 2, 1, 3).reshape(C, H * F).contiguous()
        self.a = torch.empty((self.N, 8), dtype=torch.float32, device=dx.device)
        self.alpha = torch.empty((max(self.csr.nnz, 1), H), dtype=torch.float32, device=dx.device)
        self.H = H

    def project(self, r0, r1):
        self.a[r0:r1] = ops.rowdot8(self.dx[r0:r1], self.v)

    def finish(self, r0, r1):
        L = self.layer
        if 1 <= self.g.max_degree("sl") <= 8:                   # softmax inside the fused kernel
            return ops.gatw_gemm_sm(self.dx, self.a, self.csr.rowptr[r0:r1 + 1], self.csr.col, self.wp, L.bias, r1 - r0, self.H,
                                    L.negative_slope, 0.0, 0, False, self.g.max_degree("sl"), band=0, row0=r0)[0]
        ops.gat_alpha(self.a, self.csr.rowptr, self.csr.col, self.H, L.negative_slope, 0.0, 0, False, rows=(r0, r1),
                      alpha_out=self.alpha)
        return ops.gatw_gemm(self.dx, self.csr.rowptr[r0:r1 + 1], self.csr.col, None, self.alpha, self.wp, L.bias, r1 - r0,
                             self.H, band=0)


class _TConvStages(_Stages):
    variant = "raw"

    def __init__(self, layer, dx, g):
        super().__init__(layer, dx, g)
        L = layer
        if not L._aggregate_first(dx) or L.lin_edge is not None:
            raise NotImplementedError
        self.mq, self.cq, self.w_out, self.b_out = _cached_fold(L, 'folded', L._fold_params(), dx.dtype, lambda: L._folded(dx.dtype))
        self.H = L.heads
        self.u = torch.empty((self.N, self.H * dx.shape[1]), dtype=dx.dtype, device=dx.device)

    def project(self, r0, r1):
        ops.linear_fwd(self.dx[r0:r1], self.mq, self.cq, out=self.u[r0:r1])

    def finish(self, r0, r1):
        z_aug, _ = ops.tz_fwd(self.dx, self.u[r0:r1], self.H, self.csr.rowptr[r0:r1 + 1], self.csr.col, 0.0, 0, False, band=0,
                              x_self=self.dx[r0:r1])
        out, _ = ops.linear_fwd(z_aug, self.w_out, self.b_out)
        return out


@torch.no_grad()
def forward_host(layer, x_host: torch.Tensor, edge_index_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                 rows_per_chunk: int = 1 << 18, partition=None) -> torch.Tensor:
    """`layer(x, edge_index)` (gnn_model.py:166-170) for host-resident x [N, F] / edge_index [2, E] and a host result, for
    GCNConv / GATConv / GINConv / TransformerConv in eval mode: the reference's `batch.to(device)` ... `.cpu()` round trip
    (train.py:167, inference.py:87) with the host->device copy, the kernels and the device->host copy pipelined over row
    chunks (see the module docstring; GCNConv: `gcn_forward_host`).  Inference only.  Layer configurations without a
    row-range pipeline (project-first attention paths, edge features, fp32 GAT) copy everything, run the layer and copy
    the result back — still through pinned buffers on the copy streams."""
    if isinstance(layer, GCNConv):
        return gcn_forward_host(layer, x_host, edge_index_host, out_host, rows_per_chunk, partition)
    if partition is not None and partition.world > 1:
        raise NotImplementedError("b2g.streaming: the partitioned host pipeline exists for GCNConv")
    if not isinstance(layer, (GATConv, GINConv, TransformerConv)):
        raise NotImplementedError(f"b2g.streaming: no host pipeline for {type(layer).__name__}")
    if layer.training:
        raise RuntimeError("b2g.streaming: inference only (layer.eval())")
    par = next(layer.parameters())
    if not par.is_cuda:
        raise RuntimeError("b2g.streaming: the layer must live on a CUDA device (no CPU fallback)")
    dev = par.device
    if x_host.dim() != 2 or edge_index_host.dim() != 2 or edge_index_host.shape[0] != 2:
        raise ValueError("x must be [N,F] and edge_index [2,E]")
    if x_host.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"b2g: x dtype {x_host.dtype} not supported (float32 / bfloat16)")
    hx = _pinned(x_host.contiguous(), "x")
    hei = _pinned(edge_index_host.long().contiguous(), "edge_index")
    N = hx.shape[0]
    F_out = layer.nn[-1].out_features if isinstance(layer, GINConv) else layer.out_channels * (layer.heads if layer.concat else 1)
    if out_host is None:
        out_host = torch.empty((N, F_out), dtype=hx.dtype).pin_memory()
    elif out_host.is_cuda or not out_host.is_pinned() or out_host.shape != (N, F_out) or out_host.dtype != hx.dtype:
        raise ValueError("out_host must be a pinned host tensor [N, out_channels] of x's dtype")
    if N == 0:
        return out_host
    s_in, s_cmp, s_out = _streams(dev)
    cur = torch.cuda.current_stream(dev)
    for st in (s_in, s_cmp, s_out):
        st.wait_stream(cur)
    R = max(1024, int(rows_per_chunk))
    bounds = [(r0, min(r0 + R, N)) for r0 in range(0, N, R)]
    with torch.cuda.stream(s_in):
        dei = torch.empty(hei.shape, dtype=torch.int64, device=dev)
        dei.copy_(hei, non_blocking=True)
        ev_ei = torch.cuda.Event()
        ev_ei.record(s_in)
        dx = torch.empty(tuple(hx.shape), dtype=hx.dtype, device=dev)
        ev_x = []
        for r0, r1 in bounds:
            dx[r0:r1].copy_(hx[r0:r1], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_in)
            ev_x.append(e)
    with torch.cuda.stream(s_cmp):
        for t in (dei, dx):
            t.record_stream(s_cmp)
        s_cmp.wait_event(ev_ei)
        g = Graph(dei, N)
        cls = _GINStages if isinstance(layer, GINConv) else (_GATStages if isinstance(layer, GATConv) else _TConvStages)
        try:
            stages = cls(layer, dx, g)
        except NotImplementedError:
            stages = None
        if stages is None:                      # no row-range pipeline for this configuration: whole-graph call
            s_cmp.wait_event(ev_x[-1])
            out = layer(dx, dei)
            out.record_stream(s_out)
            e = torch.cuda.Event()
            e.record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                out_host.copy_(out, non_blocking=True)
        else:
            band = int((dei[0] - dei[1]).abs().max()) if dei.shape[1] else 0          # one reduction + host sync; copies are queued
            done_rows, pending = 0, list(range(len(bounds)))

            def finish_ready():
                for c in list(pending):
                    r0, r1 = bounds[c]
                    if min(r1 + band, N) > done_rows:
                        continue
                    o = stages.finish(r0, r1)
                    o.record_stream(s_out)
                    e = torch.cuda.Event()
                    e.record(s_cmp)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(e)
                        out_host[r0:r1].copy_(o, non_blocking=True)
                    pending.remove(c)

            for c, (r0, r1) in enumerate(bounds):
                s_cmp.wait_event(ev_x[c])
                stages.project(r0, r1)
                done_rows = r1
                finish_ready()
            finish_ready()
    cur.wait_stream(s_out)
    cur.wait_stream(s_cmp)
    s_out.synchronize()
    return out_host
