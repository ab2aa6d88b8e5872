#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric   : message-passing edges/sec (whole job) = edges one GCNConv layer aggregates / step time
workload : cfg4 — synthetic 3-D hex mesh 250x200x200 = 10 M cells PER GPU (lexicographic ids,
           OpenFOAM owner/neighbour face list -> device builder mode A -> edge_index), GCNConv(256,256)
           layer forward = K6 Linear + K2 fused-normalisation CSR segment-sum + bias.  E_sl = 69.72 M
           aggregated edges per 10 M cells.  Inputs are > 5 GB, far larger than the 126 MB L2.
step     : one pass of the hot path = `GCNConv.forward(x, edge_index)` over the whole mesh.
value    : device-resident inputs, CSR cached (static mesh), CUDA-event timed, max over ranks.
e2e      : the same layer call from HOST buffers: pinned x and edge_index are copied host->device, the
           CSR is rebuilt (a new edge_index tensor every step, as `batch.to(device)` does at
           train.py:167), the layer runs, the output is copied device->host; all inside the timed region.
roofline : the dominant kernel (seg_sum_rows_kernel, K2), algorithmic bytes 2*N*F*s + 4*nnz + 4*(N+1) + 4*N
            / its own CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference : the reference's CPU path for this layer = the fp32 pure-torch
           restatement of PyG's GCNConv (oracle/layers_oracle.py; PyG itself is not installable here),
           all host threads, on a bounded sample of the same workload (a smaller hex block).
N > 1    : weak scaling — every rank owns a 250x200x200 block of a 250x200x(200 N) mesh (what RCB gives
           on this domain), with a real per-layer halo exchange of boundary rows over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, NZ = 250, 200, 200
F = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2g", choices=["b2g", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--layer", default="GCN", choices=["GCN", "GAT", "GIN", "Transformer"])
    ap.add_argument("--small", action="store_true", help="64^3 mesh: for ncu captures and CPU-side debugging")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-layer-type / train-step extras")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--train-step", action="store_true",
                    help="also time the partitioned FlowGNN train step (GAT L=6 F=256 bf16, cfg5 shape, weak scaling)")
    ap.add_argument("--train-cells", type=int, default=2500000, help="cells per GPU for --train-step")
    ap.add_argument("--train-checkpoint", action="store_true", help="re-run each layer block in backward (memory)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [t.strip() for t in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU reference
def cpu_reference(layer, steps, warmup, sample=(80, 80, 80)):
    """The reference's CPU path for one layer forward: fp32 pure-torch restatement of PyG (oracle port)."""
    import torch
    from oracle import builder_oracle as bo
    from oracle import layers_oracle as lo
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nx, ny, nz = sample
    o, n = hex_mesh_faces(nx, ny, nz)
    N = nx * ny * nz
    ei = torch.from_numpy(bo.build_graph(dict(owner=o.numpy(), neighbour=n.numpy(), cell_centers=torch.zeros(N, 3).numpy(),
                                              n_cells=N), filter_internal=True, n_internal_cells=N)['edge_index'])
    torch.manual_seed(0)
    x = torch.randn(N, F)
    H = 4
    g = lambda *s: lo.glorot_(torch.empty(*s))
    if layer == "GCN":
        W, b = g(F, F), torch.zeros(F)
        fn = lambda: lo.gcn_conv(x, ei, W, b)
        e_agg = int((ei[0] != ei[1]).sum()) + N
    elif layer == "GAT":
        W, a_s, a_d, b = g(H * F, F), g(1, H, F), g(1, H, F), torch.zeros(F)
        fn = lambda: lo.gat_conv(x, ei, W, a_s, a_d, b, heads=H)
        e_agg = int((ei[0] != ei[1]).sum()) + N
    elif layer == "GIN":
        w1, w2, b1, b2 = g(F, F), g(F, F), torch.zeros(F), torch.zeros(F)
        fn = lambda: lo.gin_conv(x, ei, lo.gin_mlp(w1, b1, w2, b2))
        e_agg = ei.shape[1]
    else:
        ws = [g(H * F, F) for _ in range(3)] + [g(F, F)]
        bs = [torch.zeros(H * F) for _ in range(3)] + [torch.zeros(F)]
        fn = lambda: lo.transformer_conv(x, ei, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2], ws[3], bs[3], heads=H)
        e_agg = ei.shape[1]
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=e_agg / dt, unit="edges/s", cores=cores, kind="port",
                sample=f"hex {nx}x{ny}x{nz} ({N} cells, {e_agg} aggregated edges) {layer}Conv F={F} fp32 forward, "
                       f"{steps} timed steps, torch CPU ops PyG dispatches to (oracle/layers_oracle.py)"), dt * 1e3, e_agg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms, e_agg = cpu_reference(args.layer, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": "message_passing_edges_per_sec", "value": cb["value"], "unit": "edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg4 hex mesh GCNConv({F},{F}) layer forward; CPU arm timed on a bounded sample: "
                                   + cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))



# ------------------------------------------------------------------------------------ parity of the timed outputs
def sampled_parity(kind, layer, x, ei, n_nodes, out, plane, tol, count=4096, edge_attr=None):
    """Checker, outside every timed region: `count` target rows of the output the timed run produced (first / last rows,
    rows next to multiples of `plane` = block faces, random interior rows) against oracle/layers_oracle.py in fp64 on the
    rows' one-hop closure (oracle/sampled.py), from the same weights and inputs.  max |mine - ref| / max |ref|."""
    import torch
    from oracle import sampled
    rows = sampled.pick_rows(out.shape[0], count, plane=plane, seed=0)
    nodes, ei_sub, pos = sampled.closure_subgraph(ei, rows.to(ei.device), n_nodes)
    ref = sampled.layer_rows(kind, layer.state_dict(), x[nodes], ei_sub, pos)
    mine = out[rows.to(out.device)].double().cpu()
    err = float((mine - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    return {"rows": int(rows.numel()), "closure_nodes": int(nodes.numel()), "closure_edges": int(ei_sub.shape[1]),
            "max_rel": err, "tol": tol, "ok": bool(err < tol),
            "oracle": "oracle/layers_oracle.py fp64 on the one-hop closure of the sampled rows (oracle/sampled.py)"}


def halo_check(b2g, layer, layer_name, world, rank, dev, dtype, tol):
    """N > 1 checker: a small mesh (64 x 64 x 16 N) through the SAME slab_partition_hex / wrap_forward path (halo exchange
    over NCCL included) against a monolithic run of the whole small mesh on this rank's GPU; every rank compares its owned
    rows, and separately the two planes next to its slab cuts (the rows that read ghost rows)."""
    import torch
    import torch.distributed as dist
    from gnn_bfs_rans_b200.distributed import slab_partition_hex
    sx, sy, sz = 64, 64, 16
    n0 = sx * sy * sz
    with torch.no_grad():
        ps = slab_partition_hex(sx, sy, sz, world, rank, dev)
        g = torch.Generator().manual_seed(4242)
        xg = torch.randn(n0 * world, F, generator=g).to(dtype).to(dev)          # the same global features on every rank
        xl = torch.zeros(ps.n_local, F, device=dev, dtype=dtype)
        xl[:n0] = xg[rank * n0:(rank + 1) * n0]
        out_p = ps.wrap_forward(layer)(xl, ps.edge_index)[:n0].float()
        mono = slab_partition_hex(sx, sy, sz * world, 1, 0, dev)
        out_m = mono.wrap_forward(layer)(xg, mono.edge_index)[rank * n0:(rank + 1) * n0].float()
        scale = out_m.abs().max().clamp_min(1e-30)
        d = (out_p - out_m).abs()
        plane = sx * sy
        cut = torch.cat([d[:plane], d[-plane:]])
        res = torch.tensor([float(d.max() / scale), float(cut.max() / scale), float((d.max(1).values > 0).sum()), float(n0)],
                           device=dev, dtype=torch.float64)
        allr = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(allr, res)
        allr = torch.stack(allr).cpu()
    mx, mxc = float(allr[:, 0].max()), float(allr[:, 1].max())
    return {"mesh": f"{sx}x{sy}x{sz * world} hex, {world} slabs", "layer": layer_name, "rows_per_rank": n0,
            "max_rel_all_rows": mx, "max_rel_cut_planes": mxc, "rows_not_bit_equal": int(allr[:, 2].sum()), "tol": tol,
            "ok": bool(mx < tol), "compares": "partitioned (wrap_forward + NCCL halo exchange) vs monolithic, same kernels"}

def run_strong_scaling(b2g, ops, layer, layer_name, world, rank, dev, dtype, timed, steps):
    """BASELINE cfg4 as written: ONE 250x200x200 (10 M-cell) mesh cut by recursive coordinate bisection of the cell centres
    into `world` parts (2x2x2 blocks of 125x100x100 at 8), general partitioner (distributed.rcb_partition + build_partition:
    every rank derives its share and the halo plan from the global edge list, no communication), one halo exchange per
    layer over NCCL.  Fixed total work: "scaling": "strong"."""
    import torch
    import torch.distributed as dist
    from gnn_bfs_rans_b200.distributed import build_partition, rcb_partition
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    ng = NX * NY * NZ
    o, n = hex_mesh_faces(NX, NY, NZ, device=dev)
    ei = ops.build_graph_edges(o, n, 1, None, ng, ng)
    del o, n
    ids = torch.arange(ng, device=dev)
    centers = torch.stack([(ids % NX).float(), ((ids // NX) % NY).float(), (ids // (NX * NY)).float()], dim=1) + 0.5
    del ids
    t0 = time.perf_counter()
    pv = rcb_partition(centers, world)
    part = build_partition(ei, pv, rank, world, device=dev)
    torch.cuda.synchronize()
    t_part = time.perf_counter() - t0
    del ei, pv, centers
    torch.cuda.empty_cache()
    torch.manual_seed(4321 + rank)
    x = torch.randn(part.n_local, F, device=dev).to(dtype)
    fwd = part.wrap_forward(layer)
    ms = timed(lambda: fwd(x, part.edge_index), steps, 3)
    t = torch.tensor([part.aggregated_edges(layer_name), part.n_owned, part.n_ghost], device=dev, dtype=torch.int64)
    mx = t.clone()
    dist.all_reduce(t)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return {"scaling": "strong", "mesh": f"{NX}x{NY}x{NZ} hex = {ng} cells in total", "partition": f"RCB into {world} parts "
            "(distributed.rcb_partition + build_partition), 1-ring halo exchange per layer (NCCL)", "n_gpus": world,
            "layer": f"{layer_name}Conv({F},{F}) forward", "cells_per_gpu_max": int(mx[1]), "ghost_rows_per_gpu_max": int(mx[2]),
            "edges_per_step": int(t[0]), "ms_per_step": ms, "edges_per_sec": int(t[0]) / (ms * 1e-3),
            "partition_build_s": t_part}


# ------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import _lib, ops
    from gnn_bfs_rans_b200.graph import Graph
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    _lib.load()  # fail loudly if the CUDA extension is missing

    nx, ny, nz = (64, 64, 64) if args.small else (NX, NY, NZ)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    s = 2 if args.dtype == "bf16" else 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # ---- mesh -> edge_index through the device builder (mode A), per rank block + halo plan
    from gnn_bfs_rans_b200.distributed import slab_partition_hex
    part = slab_partition_hex(nx, ny, nz, world, rank, dev)      # owned block + ghost planes + exchange plan
    N = part.n_owned
    ei = part.edge_index                                           # local ids, ghosts >= N
    torch.manual_seed(1234 + rank)
    x = torch.randn(part.n_local, F, device=dev).to(dtype)      # owned rows first, ghost rows (filled by the halo exchange) after
    layer = {"GCN": lambda: b2g.nn.GCNConv(F, F),
             "GAT": lambda: b2g.nn.GATConv(F, F, heads=4, concat=False),
             "GIN": lambda: b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F))),
             "Transformer": lambda: b2g.nn.TransformerConv(F, F, heads=4, concat=False)}[args.layer]()
    torch.manual_seed(0)
    layer.reset_parameters()
    layer = layer.to(dev).to(dtype).eval()

    fwd = part.wrap_forward(layer)                                 # adds the halo exchange when world > 1
    e_agg_local = part.aggregated_edges(args.layer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- value: device-resident, CSR cached
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _lib.launch_count_reset()
    ms = timed(lambda: fwd(x, ei), args.steps, max(args.warmup, 3))
    launches = _lib.launch_count()
    clk = clocks.stop() if rank == 0 else None
    launches_per_step = launches / (args.steps + max(args.warmup, 3))
    e_total = e_agg_local
    if world > 1:
        t = torch.tensor([e_agg_local], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        e_total = int(t)
    value = e_total / (ms * 1e-3)

    # ---- parity of what was just timed (checker, outside the timed region): sampled rows vs the fp64 oracle at N = 1,
    # partitioned vs monolithic on a small mesh through the same exchange path at N > 1
    tol = 2e-2 if args.dtype == "bf16" else 1e-5
    parity = halo = None
    try:
        with torch.no_grad():
            if world == 1:
                parity = sampled_parity(args.layer, layer, x, ei, part.n_local, fwd(x, ei), nx * ny, tol)
            else:
                halo = halo_check(b2g, layer, args.layer, world, rank, dev, dtype, tol)
    except Exception as e:  # the checker must not take the measurement down; a failure is reported as such
        parity = {"error": str(e)[:200], "ok": False} if world == 1 else None
        halo = {"error": str(e)[:200], "ok": False} if world > 1 else None
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (K2 seg_sum for GCN/GIN; fused attention for GAT/Transformer)
    g = b2g.graph.graph_of(ei, part.n_local)
    roof = None
    with torch.no_grad():
        if args.layer in ("GCN", "GIN"):
            variant = "sl" if args.layer == "GCN" else "raw"
            csr = g.csr(variant, False)
            dinv = g.dinv() if args.layer == "GCN" else None
            xin = torch.randn(part.n_local, F, device=dev).to(dtype)
            out = torch.empty(N, F, device=dev, dtype=dtype)
            kbias = layer.bias.float() if args.layer == "GCN" and layer.bias is not None else None
            kfn = lambda: ops.seg_sum(xin, csr.rowptr, csr.col, N, dinv, None, 0.0 if args.layer == "GCN" else 1.0, None,
                                      kbias, out=out, band=g.band())   # the launch the layer forward makes (same template instance)
            kms = timed(kfn, args.steps, 3)
            alg = 2 * N * F * s + 4 * csr.nnz + 4 * (N + 1) + (4 * N if dinv is not None else 0)
            kname = "seg_rows_kernel (K2/K3, aggregate_rows.cu)"
            gather_row_bytes = F * s
        else:
            H = 4
            kms, alg, kname = None, None, "attn_fwd_kernel (K4/K5)"
            if args.layer == "GAT":
                csr = g.csr("sl", False)
                xin = torch.randn(part.n_local, F, device=dev).to(dtype)
                if ops.gatw_gemm_supported(N, H, F, F, dtype):
                    # fused aggregation + projection (gat_fused.cu): reads x once, col + alpha [nnz, 4]; writes out [N, C]
                    alpha = torch.rand(max(csr.nnz, 1), H, device=dev)
                    wp = torch.randn(F, H * F, device=dev).to(dtype) / 16
                    obuf = torch.empty(N, F, device=dev, dtype=dtype)
                    kfn = lambda: ops.gatw_gemm(xin, csr.rowptr, csr.col, None, alpha, wp, None, N, H, band=g.band(), out=obuf)
                    alg = N * F * s + N * F * s + 16 * csr.nnz + 4 * csr.nnz + 4 * (N + 1)
                    kname = "gatw_gemm_kernel (K4f fused aggregation + projection, gat_fused.cu)"
                else:
                    a = torch.randn(part.n_local, 2 * H, device=dev)
                    zbuf = torch.empty(N, H * F, device=dev, dtype=dtype)
                    kfn = lambda: ops.gatz_fwd(xin, a, H, 0.2, csr.rowptr, csr.col, 0.0, 0, False, band=g.band(), out=zbuf)
                    # aggregate-first kernel: read x once, a [N,2H] fp32, indices; write z [N, H*F]
                    alg = N * F * s + N * H * F * s + 2 * 4 * N * H + 4 * csr.nnz + 4 * (N + 1)
                    kname = "gatz_fwd_kernel (K4 aggregate-first, gat_rows.cu)"
                gather_row_bytes = F * s
            else:
                csr = g.csr("raw", False)
                y = torch.randn(part.n_local, 3 * H * F + F, device=dev).to(dtype)
                q, k, v, sk = y[:, :H * F], y[:, H * F:2 * H * F], y[:, 2 * H * F:3 * H * F], y[:, 3 * H * F:]
                kfn = lambda: ops.tconv_fwd(q, k, v, sk, H, F, False, csr.rowptr, csr.col, 0.0, 0, False)
                alg = 3 * N * H * F * s + 2 * N * F * s + 4 * csr.nnz + 4 * (N + 1)
                gather_row_bytes = 2 * H * F * s                       # k and v rows per edge
            kms = timed(kfn, args.steps, 3)
        ach = alg / (kms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "algorithmic_bytes": alg, "kernel_ms": kms,
                "kernel_edges_per_sec": e_agg_local / (kms * 1e-3), "peak_source": peak_src,
                # secondary figure: bytes the gathers request, E * row bytes; L2 serves the re-reads, so it may
                # exceed the HBM peak and is never the roofline fraction
                "gather_model_gb_s": int(csr.nnz) * gather_row_bytes / (kms * 1e-3) / 1e9}
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            try:
                tj = json.load(open(tfile))
                roof["traffic"] = tj.get(f"{args.layer}_{args.dtype}")
                roof["traffic_source"] = (tj.get("source") or {}).get(f"{args.layer}_{args.dtype}")
            except Exception:
                pass

    # ---- e2e: host buffers in, host buffer out, CSR rebuilt each step (fresh edge_index tensor), at N GPUs.
    # N = 1: the public host-buffer entry with the copies pipelined against the kernels (streaming.py).
    # N > 1: every rank moves its own share (pinned host -> device), rebuilds its local CSR (+ the ghost deg^-1/2
    #        exchange), runs the layer with the halo exchange and copies its rows back; wall clock, max over ranks.
    e2e = None

    def host_mem_ok(need_bytes):
        try:
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024 > 1.5 * need_bytes
        except Exception:
            pass
        return True

    numa_bound = b2g.streaming.bind_host_to_gpu_numa(dev) if world > 1 else False   # pinned buffers on the GPU's NUMA node
    need = (N * F * (2 if args.dtype == "bf16" else 4) * 2 + ei.numel() * 8) * world
    if not host_mem_ok(need):
        e2e = {"value": None, "unit": "edges/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "note": f"skipped: {need / 1e9:.0f} GB of pinned host buffers would not fit this box's free memory"}
    else:
        with torch.no_grad():
            hx = torch.empty((N, F), dtype=dtype, pin_memory=True)
            hx.copy_(x[:N])
            hei = torch.empty(tuple(ei.shape), dtype=torch.int64, pin_memory=True)
            hei.copy_(ei)
            hout = torch.empty((N, F), dtype=dtype, pin_memory=True)

            if args.layer == "GCN":
                def e2e_step():
                    b2g.streaming.gcn_forward_host(layer, hx, hei, hout, partition=part if world > 1 else None)
                e2e_api = ("gnn_bfs_rans_b200.streaming.gcn_forward_host(layer, x_host, edge_index_host, out_host"
                           + (", partition=...) on every rank (one halo exchange of the projected boundary rows)" if world > 1 else ")"))
                ei_keep = part.edge_index
            elif world == 1:
                def e2e_step():
                    b2g.streaming.forward_host(layer, hx, hei, hout)
                e2e_api = "gnn_bfs_rans_b200.streaming.forward_host(layer, x_host, edge_index_host, out_host)"
            else:
                ei_keep = part.edge_index

                def e2e_step():
                    dx = hx.to(dev, non_blocking=True)
                    dei = hei.to(dev, non_blocking=True)
                    part.edge_index, part._dinv_ready = dei, False      # a fresh edge_index: local CSR + ghost dinv rebuilt
                    o = part.wrap_forward(layer)(dx, dei)
                    hout.copy_(o, non_blocking=True)
                e2e_api = ("per rank: x.to(dev), edge_index.to(dev), Partition.wrap_forward(layer) (CSR rebuild, halo "
                           "exchange over NCCL), pinned host copy of the owned rows")

            steps_e = max(3, min(args.steps, 5))
            for _ in range(2):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps_e):
                e2e_step()
            e1.record()
            barrier()
            wall = (time.perf_counter() - t0) / steps_e
            ems = max(e0.elapsed_time(e1) / steps_e, wall * 1e3)
            if world > 1:
                t = torch.tensor([ems], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ems = float(t)
                part.edge_index, part._dinv_ready = ei_keep, False
                part.prepare_graph()
            e2e = {"value": e_total / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems,
                   "h2d_bytes_per_step": (hx.numel() * hx.element_size() + hei.numel() * 8) * world,
                   "d2h_bytes_per_step": hout.numel() * hout.element_size() * world, "steps": steps_e,
                   "includes": "H2D x + edge_index, CSR rebuild, layer forward, D2H output", "api": e2e_api,
                   "host_numa_bound": bool(numa_bound)}
            del hx, hei, hout

    # ---- extras: other layer types / fp32 / fwd+bwd / FlowGNN train step (not the headline)
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = run_extras(b2g, ops, part, dev, timed)
        extras.update(run_mesh_ingest(b2g, dev, timed, (nx, ny, nz)))

    # ---- BASELINE cfg4 as a strong-scaling problem (one 10 M-cell mesh, RCB), N > 1
    strong = None
    if world > 1 and not args.no_extras:
        try:
            del x
            torch.cuda.empty_cache()
            strong = run_strong_scaling(b2g, ops, layer, args.layer, world, rank, dev, dtype, timed, args.steps)
        except Exception as e:
            strong = {"error": str(e)[:200]}
        x = None

    # ---- cfg5-shaped train step, partitioned (halo exchange per layer, synchronised BatchNorm, gradient all-reduce):
    # part of the default line at N = 8 (BASELINE cfg5: 100 M cells = 12.5 M per GPU, GAT L = 6, F = 256, bf16); opt-in elsewhere
    train_line = None
    if world == 8 and not args.no_extras and not args.train_step:
        args.train_step, args.train_checkpoint, args.train_cells = True, True, 12500000
    if args.train_step:
        x = None
        try:
            train_line = run_partitioned_train_step(args, b2g, world, rank, dev, F, nx, ny, barrier)
        except Exception as e:                                   # never lose the headline line to the extra
            train_line = {"error": str(e)[:300]}

    # ---- CPU baseline (rank 0, N=1, bounded sample)
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _, _ = cpu_reference(args.layer, 3, 1, sample=(100, 100, 100))

    if rank == 0:
        line = {"metric": "message_passing_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"cfg4: 3-D hex mesh {nx}x{ny}x{nz} = {N} cells per GPU "
                                       f"({nx}x{ny}x{nz * world} total), {args.layer}Conv({F},{F}) layer forward, "
                                       f"{e_total} aggregated edges/step",
                           "cells_per_gpu": N, "edges_per_step": e_total, "hidden": F, "layer": args.layer,
                           "partition": "none" if world == 1 else f"RCB slabs x{world}, 1-ring halo exchange per layer (NCCL)",
                           "cache_policy": "inputs (>5 GB) larger than the 126 MB L2; CSR cached across steps in `value`"},
                "parity_check": parity, "halo_check": halo, "clocks": clk, "e2e": e2e, "gpu_launches": int(round(launches_per_step * args.steps)),
                "gpu_launches_per_step": launches_per_step, "roofline": roof, "cpu_baseline": cb, "extras": extras}
        if strong is not None:
            line["strong_scaling_cfg4"] = strong
        if train_line is not None:
            line["train_step_partitioned"] = train_line
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_partitioned_train_step(args, b2g, world, rank, dev, F, nx, ny, barrier):
    """cfg5-shaped train step on this rank's slab (halo exchange per layer, synchronised BatchNorm, gradient all-reduce)."""
    import torch
    import torch.distributed as dist
    from gnn_bfs_rans_b200.distributed import slab_partition_hex
    from gnn_bfs_rans_b200.distributed import flow_forward_partitioned, allreduce_gradients
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    torch.cuda.empty_cache()
    nzs = max(2, args.train_cells // (nx * ny))
    tpart = slab_partition_hex(nx, ny, nzs, world, rank, dev)
    n_glob = tpart.n_owned * world
    torch.manual_seed(0)
    tmodel = FlowGNN(3, F, 7, 6, "GAT", dropout=0.1).to(dev).to(torch.bfloat16).train()
    topt = torch.optim.Adam(tmodel.parameters(), lr=3e-4, weight_decay=1e-5)
    txin = torch.rand(tpart.n_owned, 3, device=dev, dtype=torch.bfloat16)
    ty = torch.rand(tpart.n_owned, 7, device=dev, dtype=torch.bfloat16)

    def tstep():
        topt.zero_grad(set_to_none=True)
        o = flow_forward_partitioned(tmodel, txin, tpart, checkpoint_layers=args.train_checkpoint)
        loss = (o - ty).float().square().sum() / (n_glob * 7)
        loss.backward()
        allreduce_gradients(list(tmodel.parameters()), world)
        torch.nn.utils.clip_grad_norm_(tmodel.parameters(), 1.0)
        topt.step()
    for _ in range(2):
        tstep()
    barrier()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        tstep()
    e1.record()
    barrier()
    tms = e0.elapsed_time(e1) / 3
    if world > 1:
        t = torch.tensor([tms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tms = float(t)
    train_line = {"model": "FlowGNN GAT L=6 F=256 bf16 (cfg5 shape), fwd + loss + bwd + gradient all-reduce + clip + Adam",
                  "cells_per_gpu": tpart.n_owned, "cells_total": n_glob, "ms_per_step": tms,
                  "cells_per_sec": n_glob / (tms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                  "recompute": os.environ.get("B2G_RECOMPUTE", "0"), "checkpoint_layers": bool(args.train_checkpoint)}
    del tmodel, topt
    return train_line



def run_mesh_ingest(b2g, dev, timed, dims):
    """Not the headline: the device mesh ingest (csrc/mesh.cu) on the bench mesh as a polyMesh — cell
    centres + internal-cell mask + n_cells from points / faces / owner / neighbour already in HBM — and the numpy oracle
    of the same computation on a bounded sample (the reference itself loops over every face in Python)."""
    import time
    import torch
    from gnn_bfs_rans_b200.synthetic import hex_cell_centers, hex_polymesh
    out = {}
    try:
        nx, ny, nz = dims
        pts, own, nbr, fp, fo = hex_polymesh(nx, ny, nz, dev)
        res = {}

        def run():
            res["d"] = b2g.mesh.derive_mesh(pts, own, nbr, (fp, fo), dev, as_numpy=False)
        ms = timed(run, 5, 3)
        d = res["d"]
        n_cells, slots = d["n_cells"], int(fo[own.numel()]) + int(fo[nbr.numel()])
        # K0m algorithmic bytes: ids + offsets read twice, vertex lists written + read, unique vertices, output
        alg = 2 * (4 * (own.numel() + nbr.numel()) + 16 * (own.numel() + nbr.numel()) + 4 * slots) + 8 * slots \
            + 24 * 8 * n_cells + 24 * n_cells + 2 * 4 * nbr.numel() + n_cells
        ok = bool(torch.equal(d["cell_centers"][:nx * ny].cpu(), torch.from_numpy(hex_cell_centers(nx, ny, 1))))
        out["mesh_ingest"] = {"cells": n_cells, "faces": int(own.numel()), "vertex_slots": slots, "ms": ms,
                              "cells_per_sec": n_cells / (ms * 1e-3), "algorithmic_gb_s": alg / (ms * 1e-3) / 1e9,
                              "first_plane_exact": ok, "all_internal": bool(d["internal_mask"].all())}
        del pts, own, nbr, fp, fo, d, res
        torch.cuda.empty_cache()
        from oracle import mesh_oracle as mo                      # the checker, timed as the CPU baseline of this extra
        sx, sy, sz = min(nx, 100), min(ny, 100), min(nz, 50)
        p2, o2, n2, fp2, fo2 = [t.numpy() for t in hex_polymesh(sx, sy, sz)]
        t0 = time.perf_counter()
        mo.get_cell_centers(p2, o2, n2, fp2, fo2)
        mo.get_internal_cells(o2, n2)
        dt = time.perf_counter() - t0
        out["mesh_ingest"]["cpu_oracle"] = {"sample": f"{sx}x{sy}x{sz} hex block, numpy restatement, 1 thread",
                                            "cells_per_sec": sx * sy * sz / dt}
    except Exception as e:
        out["mesh_ingest"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()
    return out


def run_extras(b2g, ops, part, dev, timed):
    """Not the headline: per-layer-type forward and forward+backward throughput, and the reference's train step
    (FlowGNN forward + MSE loss + backward + clip_grad_norm_ + Adam, train.py:170-189) on the same mesh."""
    import torch
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    out = {}
    N, ei = part.n_owned, part.edge_index

    def mk(lt):
        return {"GCN": lambda: b2g.nn.GCNConv(F, F),
                "GAT": lambda: b2g.nn.GATConv(F, F, heads=4, concat=False),
                "GIN": lambda: b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F))),
                "Transformer": lambda: b2g.nn.TransformerConv(F, F, heads=4, concat=False)}[lt]()

    def timed_grad(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for dt_name, dtype in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        for lt in ("GCN", "GAT", "GIN", "Transformer"):
            key = f"{lt}_{dt_name}"
            try:
                torch.manual_seed(0)
                layer = mk(lt).to(dev).to(dtype).eval()
                x = torch.empty(N, F, device=dev, dtype=dtype).normal_()
                ms = timed(lambda: layer(x, ei), 5, 2)
                e_agg = part.aggregated_edges(lt)
                out[key + "_fwd"] = {"ms": ms, "edges_per_sec": e_agg / (ms * 1e-3)}
                try:
                    with torch.no_grad():
                        out[key + "_fwd"]["parity_check"] = sampled_parity(
                            lt, layer, x, ei, N, layer(x, ei), NX * NY, 2e-2 if dt_name == "bf16" else 1e-5, count=2048)
                except Exception as e:
                    out[key + "_fwd"]["parity_check"] = {"error": str(e)[:200], "ok": False}
                if dt_name == "bf16":
                    xg = x.requires_grad_(True)

                    def fb():
                        xg.grad = None
                        layer.zero_grad(set_to_none=True)
                        layer(xg, ei).backward(gout)
                    gout = torch.empty(N, F, device=dev, dtype=dtype).normal_()
                    ms = timed_grad(fb, 3, 2)
                    out[key + "_fwd_bwd"] = {"ms": ms, "edges_per_sec": e_agg / (ms * 1e-3)}
                    del xg, gout
                del layer, x
            except Exception as e:  # an extra must never take the headline down
                out[key] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()

    # host-buffer (e2e) entry for the other layer types: streaming.forward_host from pinned host buffers, copies in the timed region
    try:
        hei = torch.empty(tuple(ei.shape), dtype=torch.int64, pin_memory=True)
        hei.copy_(ei)
        hx = torch.empty((N, F), dtype=torch.bfloat16, pin_memory=True).normal_()
        hout = torch.empty((N, F), dtype=torch.bfloat16, pin_memory=True)
        for lt in ("GAT", "GIN", "Transformer"):
            torch.manual_seed(0)
            layer = mk(lt).to(dev).to(torch.bfloat16).eval()
            for _ in range(2):
                b2g.streaming.forward_host(layer, hx, hei, hout)
            torch.cuda.synchronize()
            import time as _t
            t0 = _t.perf_counter()
            for _ in range(3):
                b2g.streaming.forward_host(layer, hx, hei, hout)
            torch.cuda.synchronize()
            ms = (_t.perf_counter() - t0) / 3 * 1e3
            e_agg = part.aggregated_edges(lt)
            out[f"{lt}_bf16_e2e_host"] = {"ms": ms, "edges_per_sec": e_agg / (ms * 1e-3),
                                          "h2d_bytes": hx.numel() * 2 + hei.numel() * 8, "d2h_bytes": hout.numel() * 2,
                                          "api": "streaming.forward_host"}
            del layer
        del hx, hei, hout
    except Exception as e:
        out["e2e_host_layers"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # TransformerConv(edge_dim=4): the same layer consuming [dir, dist] edge attributes of the mesh
    try:
        torch.manual_seed(0)
        layer = b2g.nn.TransformerConv(F, F, heads=4, concat=False, edge_dim=4).to(dev).to(torch.bfloat16).eval()
        x = torch.empty(N, F, device=dev, dtype=torch.bfloat16).normal_()
        ea = torch.empty(ei.shape[1], 4, device=dev, dtype=torch.float32).normal_()
        ms = timed(lambda: layer(x, ei, edge_attr=ea), 5, 2)
        e_agg = part.aggregated_edges("Transformer")
        out["Transformer_edge_dim4_bf16_fwd"] = {"ms": ms, "edges_per_sec": e_agg / (ms * 1e-3)}
        xg = x.requires_grad_(True)
        gout = torch.empty(N, F, device=dev, dtype=torch.bfloat16).normal_()

        def fbe():
            xg.grad = None
            layer.zero_grad(set_to_none=True)
            layer(xg, ei, edge_attr=ea).backward(gout)
        ms = timed_grad(fbe, 3, 2)
        out["Transformer_edge_dim4_bf16_fwd_bwd"] = {"ms": ms, "edges_per_sec": e_agg / (ms * 1e-3)}
        del layer, x, xg, gout, ea
    except Exception as e:
        out["Transformer_edge_dim4_bf16"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # FlowGNN train step (fwd + loss + bwd + clip + Adam; hidden 256, 4 layers, bf16, train.py:170-189) on a 2.5 M-cell
    # block of the same mesh: the reference caller as is (drop-in layers + BatchNorm) and with the glue fused
    # (FlowGNN(fused_glue=True): residual + BatchNorm + ReLU + dropout in the two passes of csrc/bn.cu)
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    n_sub = N // 4
    o, nb = hex_mesh_faces(NX, NY, NZ // 4, device=dev)
    sub_ei = ops.build_graph_edges(o, nb, 1, None, n_sub, n_sub)
    for lt in ("GCN", "GAT"):
        for fused in (False, True):
            key = f"train_step_FlowGNN_{lt}_L4_F256_bf16" + ("_fused_glue" if fused else "")
            try:
                torch.manual_seed(0)
                torch.cuda.reset_peak_memory_stats()
                model = FlowGNN(3, F, 7, 4, lt, dropout=0.1, fused_glue=fused).to(dev).to(torch.bfloat16).train()
                opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5)
                xin = torch.rand(n_sub, 3, device=dev, dtype=torch.bfloat16)
                y = torch.rand(n_sub, 7, device=dev, dtype=torch.bfloat16)

                def step():
                    opt.zero_grad(set_to_none=True)
                    loss = (model(xin, sub_ei) - y).float().square().mean()
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                    opt.step()
                ms = timed_grad(step, 5, 2)
                out[key] = {"ms": ms, "cells": n_sub, "edges": int(sub_ei.shape[1]),
                            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
                del model, opt, xin, y
            except Exception as e:
                out[key] = {"error": str(e)[:160]}
            torch.cuda.empty_cache()
    # ---- BASELINE.json cfg2: the shipped BFS case (12 225 internal cells, mode A graph) through FlowGNN(GAT, L=4, H=4,
    # hidden 128), fp32 and bf16: inference forward and train step.  Launch-bound at this size (~100 kernels of a few us).
    try:
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "shipped_mesh.npz"))
        mesh = dict(owner=z['owner'], neighbour=z['neighbour'], cell_centers=z['cell_centers'], n_cells=int(z['n_cells']))
        gdata = b2g.GraphConstructor(mesh).build_graph(node_features=mesh['cell_centers'], filter_internal=True,
                                                       n_internal_cells=12225)
        ei2 = gdata.edge_index.to(dev)
        n2 = int(gdata.num_nodes)
        for dt_name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            for lt in ("GCN", "GAT"):
                torch.manual_seed(0)
                model = FlowGNN(3, 128, 7, 4, lt, dropout=0.1).to(dev).to(dtype)
                xin = torch.rand(n2, 3, device=dev, dtype=dtype)
                y = torch.rand(n2, 7, device=dev, dtype=dtype)
                model.eval()
                ms_f = timed(lambda: model(xin, ei2), 20, 5)
                ms_g = None
                try:                    # the same forward replayed from one CUDA graph (static mesh): launch latency removed
                    gf = b2g.graphs.GraphedForward(model, xin, ei2)
                    ms_g = timed(lambda: gf(xin), 50, 5)
                    del gf
                except Exception as e:
                    ms_g = "error: " + str(e)[:120]
                model.train()
                opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5)

                def step2():
                    opt.zero_grad(set_to_none=True)
                    loss = (model(xin, ei2) - y).float().square().mean()
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                    opt.step()
                ms_t = timed_grad(step2, 20, 5)
                ms_tg = None
                try:                    # the whole step (zero_grad, fwd, loss, bwd, clip, Adam) replayed from one CUDA graph
                    opt_c = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5, capturable=True)
                    gts = b2g.graphs.GraphedTrainStep(model, opt_c, lambda o, t: (o - t).float().square().mean(), xin, y, ei2,
                                                      max_grad_norm=1.0)
                    ms_tg = timed_grad(lambda: gts.step(xin, y), 50, 5)
                    del gts, opt_c
                except Exception as e:
                    ms_tg = "error: " + str(e)[:120]
                ms_tf = None
                if dt_name == "fp32":
                    try:                # + the reference's criterion and clip + Adam as fused kernels (training.py), fused glue
                        torch.manual_seed(0)
                        mf = FlowGNN(3, 128, 7, 4, lt, dropout=0.1, fused_glue=True).to(dev).train()
                        crit = b2g.training.WeightedMSELoss()
                        opt_f = b2g.training.FusedClipAdam(mf.parameters(), lr=3e-4, weight_decay=1e-5, max_grad_norm=1.0)
                        gtf = b2g.graphs.GraphedTrainStep(mf, opt_f, lambda o, t: crit(o, t, pressure_ref_weight=0.1), xin, y, ei2)
                        ms_tf = timed_grad(lambda: gtf.step(xin, y), 50, 5)
                        del gtf, opt_f, mf
                    except Exception as e:
                        ms_tf = "error: " + str(e)[:120]
                out[f"cfg2_shipped_BFS_FlowGNN_{lt}_L4_F128_{dt_name}"] = {"cells": n2, "edges": int(ei2.shape[1]),
                                                                           "forward_ms": ms_f, "forward_cuda_graph_ms": ms_g,
                                                                           "train_step_ms": ms_t,
                                                                           "train_step_cuda_graph_ms": ms_tg,
                                                                           "train_step_cuda_graph_fused_loss_adam_ms": ms_tf}
                del model, opt
    except Exception as e:
        out["cfg2_shipped_BFS"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # ---- BASELINE.json cfg3: 2-D unstructured mesh, ~1 M cells (Delaunay dual of 500 500 points, cell ids = triangle ids:
    # spatially incoherent order on purpose), GIN and TransformerConv hidden 256, bf16: layer forward and forward+backward
    try:
        from gnn_bfs_rans_b200.synthetic import delaunay_dual_faces, hilbert_renumber_2d
        own0, nbr0, cen0 = delaunay_dual_faces(500500, seed=1)
        n3 = int(max(own0.max(), nbr0.max())) + 1
        own1, nbr1, _, _ = hilbert_renumber_2d(own0, nbr0, cen0)      # the "Hilbert-sorted variant" of SURVEY §8d cfg3
        for tag, own, nbr in (("", own0, nbr0), ("_hilbert", own1, nbr1)):
            ei3 = ops.build_graph_edges(torch.from_numpy(own).to(dev), torch.from_numpy(nbr).to(dev), 1, None, n3, n3)
            for lt in ("GIN", "Transformer"):
                torch.manual_seed(0)
                layer = mk(lt).to(dev).to(torch.bfloat16).eval()
                x3 = torch.empty(n3, F, device=dev, dtype=torch.bfloat16).normal_()
                ms_f = timed(lambda: layer(x3, ei3), 10, 3)
                try:
                    with torch.no_grad():
                        par3 = sampled_parity(lt, layer, x3, ei3, n3, layer(x3, ei3), 0, 2e-2, count=2048)
                except Exception as e:
                    par3 = {"error": str(e)[:200], "ok": False}
                xg = x3.clone().requires_grad_(True)
                g3 = torch.empty(n3, F, device=dev, dtype=torch.bfloat16).normal_()

                def fb3():
                    xg.grad = None
                    layer.zero_grad(set_to_none=True)
                    layer(xg, ei3).backward(g3)
                ms_b = timed_grad(fb3, 5, 2)
                e3 = int(ei3.shape[1])
                out[f"cfg3_delaunay{tag}_{lt}_F256_bf16"] = {"cells": n3, "edges": e3, "forward_ms": ms_f, "fwd_bwd_ms": ms_b,
                                                             "fwd_edges_per_sec": e3 / (ms_f * 1e-3), "parity_check": par3}
                del layer, x3, xg, g3
    except Exception as e:
        out["cfg3_delaunay"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
