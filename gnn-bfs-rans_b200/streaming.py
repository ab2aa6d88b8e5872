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
from .nn import GCNConv

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
        dx = torch.empty((N, F_in), dtype=hx.dtype, device=dev)
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
        xs = torch.empty((n_local, F_out), dtype=hx.dtype, device=dev)
        out = torch.empty((N, F_out), dtype=hx.dtype, device=dev)
        out.record_stream(s_out)
        done_rows, pending = 0, list(range(len(bounds)))           # rows projected so far / chunks not yet aggregated

        def aggregate(c):
            r0, r1 = bounds[c]
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
            ops.linear_fwd(dx[r0:r1], wt, None, row_scale=dinv[r0:r1], out=xs[r0:r1])
            done_rows = r1
            aggregate_ready(False)
        if multi:
            partition.exchange(xs)             # projected (and dinv-scaled) boundary rows -> the neighbours' ghost rows
        aggregate_ready(True)
    cur.wait_stream(s_out)
    cur.wait_stream(s_cmp)
    s_out.synchronize()                        # the result is host memory: return only when it is complete
    return out_host
