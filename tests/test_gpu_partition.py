"""Multi-GPU path: partitioned == monolithic.
 * virtual ranks on ONE GPU: every rank's Partition is built in this process and the halo exchange is
   emulated by index copies between the ranks' tensors (what the driver's 1-GPU run can execute);
 * real NCCL with 2 processes when >= 2 GPUs are visible (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def make_layer(kind, F):
    import gnn_bfs_rans_b200 as b2g
    torch.manual_seed(11)
    m = {"GCN": lambda: b2g.nn.GCNConv(F, F), "GAT": lambda: b2g.nn.GATConv(F, F, heads=4, concat=False),
         "GIN": lambda: b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F))),
         "Transformer": lambda: b2g.nn.TransformerConv(F, F, heads=4, concat=False)}[kind]()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.uniform_(-0.5, 0.5)
    return m.cuda().eval()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind", ["GCN", "GAT", "GIN", "Transformer"])
def test_virtual_ranks_slab_partition(kind, world):
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.distributed import slab_partition_hex
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    nx, ny, nz, F = 12, 10, 6, 64
    # monolithic graph of the whole nx x ny x (nz*world) block
    o, n = hex_mesh_faces(nx, ny, nz * world, device='cuda')
    N = nx * ny * nz * world
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    layer = make_layer(kind, F)
    torch.manual_seed(3)
    x = torch.randn(N, F, device='cuda')
    with torch.no_grad():
        ref = layer(x, ei)
    parts = [slab_partition_hex(nx, ny, nz, world, r, 'cuda') for r in range(world)]
    nb = nx * ny * nz
    owned = [x[r * nb:(r + 1) * nb] for r in range(world)]

    def emulate_exchange(r, rows_of):
        """fill rank r's ghost rows from the owners' tensors using the PEERS' send lists"""
        def fn(buf):
            off = parts[r].n_owned
            for p in range(world):
                cnt = parts[r].recv_counts[p]
                if cnt:
                    buf[off:off + cnt] = rows_of(p)[parts[p].send_idx[r].long()]
                    off += cnt
            return buf
        return fn

    dinvs = {}
    from gnn_bfs_rans_b200.graph import graph_of
    for r in range(world):                       # local deg^-1/2 of the owned rows (complete: all in-edges present)
        dinvs[r] = graph_of(parts[r].edge_index, parts[r].n_local).dinv()[:parts[r].n_owned].clone()
    for r in range(world):
        pr = parts[r]
        pr.prepare_graph(exchange=emulate_exchange(r, lambda p: torch.stack([dinvs[p]] + [torch.zeros_like(dinvs[p])] * 3, 1)))
        xf = torch.empty(pr.n_local, F, device='cuda')
        xf[:pr.n_owned] = owned[r]
        emulate_exchange(r, lambda p: owned[p])(xf)
        with torch.no_grad():
            out = layer(xf, pr.edge_index)[:pr.n_owned]
        assert rel(out, ref[r * nb:(r + 1) * nb]) < 1e-5, (kind, world, r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.distributed import HaloFn, allreduce_gradients, slab_partition_hex
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny, nz, F = 16, 12, 8, 64
        nb = nx * ny * nz
        part = slab_partition_hex(nx, ny, nz, world, rank, f"cuda:{rank}")
        part.prepare_graph()
        layer = make_layer("GCN", F).to(f"cuda:{rank}").train()
        torch.manual_seed(3)
        x_all = torch.randn(nb * world, F, device=f"cuda:{rank}")
        xo = x_all[rank * nb:(rank + 1) * nb].clone().requires_grad_(True)
        out = layer(HaloFn.apply(xo, part), part.edge_index)[:part.n_owned]
        loss = out.square().sum()
        loss.backward()
        allreduce_gradients(list(layer.parameters()), world)
        # monolithic reference on every rank
        o, n = hex_mesh_faces(nx, ny, nz * world, device=f"cuda:{rank}")
        ei = ops.build_graph_edges(o, n, 1, None, nb * world, nb * world)
        ref_layer = make_layer("GCN", F).to(f"cuda:{rank}").train()
        xr = x_all.clone().requires_grad_(True)
        ref = ref_layer(xr, ei)
        ref.square().sum().backward()
        e_out = rel(out.detach(), ref.detach()[rank * nb:(rank + 1) * nb])
        e_gx = rel(xo.grad, xr.grad[rank * nb:(rank + 1) * nb])
        e_gw = rel(layer.lin.weight.grad, ref_layer.lin.weight.grad)
        # inference path of Partition.wrap_forward: exchange on a side stream behind the owned-row GEMM, same bits
        os.environ["B2G_HALO_OVERLAP"] = "1"
        with torch.no_grad():
            fwd = part.wrap_forward(layer)
            for _ in range(3):
                o2 = fwd(xo.detach(), part.edge_index)
            xf = torch.zeros(part.n_local, F, device=f"cuda:{rank}")
            xf[:part.n_owned] = xo.detach()
            part.exchange(xf)
            o1 = layer(xf, part.edge_index)[:part.n_owned]
        same = o2.shape == o1.shape and bool(torch.equal(o1, o2))
        q.put((rank, e_out, e_gx, e_gw, same))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_nccl_halo_forward_backward():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, e_out, e_gx, e_gw, same in res:
        assert e_out < 1e-5 and e_gx < 1e-5 and e_gw < 1e-5, res
        assert same, "overlapped halo forward (Partition.wrap_forward) differs from the blocking exchange + layer"


def _flow_worker(rank, world, port, q, layer_type, ckpt=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gnn_bfs_rans_b200 import ops
        from gnn_bfs_rans_b200.distributed import slab_partition_hex, flow_forward_partitioned, allreduce_gradients
        from gnn_bfs_rans_b200.flow_model import FlowGNN
        from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
        dev = f"cuda:{rank}"
        nx, ny, nz = 12, 10, 8
        nb = nx * ny * nz
        N = nb * world
        part = slab_partition_hex(nx, ny, nz, world, rank, dev)
        torch.manual_seed(0)
        model = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.0).to(dev).train()
        ref_model = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.0).to(dev).train()
        ref_model.load_state_dict(model.state_dict())
        torch.manual_seed(3)
        x_all = torch.rand(N, 3, device=dev)
        y_all = torch.rand(N, 7, device=dev)
        sl = slice(rank * nb, (rank + 1) * nb)
        out = flow_forward_partitioned(model, x_all[sl], part, checkpoint_layers=ckpt)
        loss = (out - y_all[sl]).square().sum() / (N * 7)              # normalised by GLOBAL counts
        loss.backward()
        allreduce_gradients(list(model.parameters()), world)
        # monolithic reference (whole mesh on this rank)
        o, n = hex_mesh_faces(nx, ny, nz * world, device=dev)
        ei = ops.build_graph_edges(o, n, 1, None, N, N)
        ref = ref_model(x_all, ei)
        ((ref - y_all).square().sum() / (N * 7)).backward()
        e_out = rel(out.detach(), ref.detach()[sl])
        e_g = 0.0
        gmax = max(float(p.grad.abs().max()) for p in ref_model.parameters())
        for (name, p), pr in zip(model.named_parameters(), ref_model.parameters()):
            e_g = max(e_g, float((p.grad - pr.grad).abs().max()) / max(float(pr.grad.abs().max()), 5e-2 * gmax))
        e_rm = max(rel(b, br) for (nm, b), br in zip(model.named_buffers(), ref_model.buffers()) if b.dtype.is_floating_point)
        q.put((rank, e_out, e_g, e_rm))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("layer_type,ckpt", [("GCN", False), ("GAT", False), ("GAT", True)])
def test_two_gpu_partitioned_flowgnn_step_equals_monolithic(layer_type, ckpt):
    """Whole model on 2 ranks (halo exchange per layer, BatchNorm statistics combined over the ranks, loss normalised by
    global counts, flat gradient all-reduce) == the single-process model on the whole mesh: outputs, every parameter
    gradient and the BatchNorm running statistics."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flow_worker, args=(r, 2, port, q, layer_type, ckpt)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, e_out, e_g, e_rm in res:
        assert e_out < 2e-5 and e_g < 2e-4 and e_rm < 1e-5, res


def _stream_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gnn_bfs_rans_b200 as b2g
        from gnn_bfs_rans_b200 import ops
        from gnn_bfs_rans_b200.distributed import slab_partition_hex
        from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
        dev = f"cuda:{rank}"
        nx, ny, nz, F = 16, 12, 40, 256
        nb = nx * ny * nz
        N = nb * world
        part = slab_partition_hex(nx, ny, nz, world, rank, dev)
        torch.manual_seed(0)
        layer = b2g.nn.GCNConv(F, F).to(dev).eval()
        with torch.no_grad():
            layer.bias.uniform_(-1, 1)
        torch.manual_seed(3)
        x_all = torch.randn(N, F)
        sl = slice(rank * nb, (rank + 1) * nb)
        out = b2g.streaming.gcn_forward_host(layer, x_all[sl], part.edge_index.cpu(), rows_per_chunk=2048, partition=part)
        o, n = hex_mesh_faces(nx, ny, nz * world, device=dev)
        ei = ops.build_graph_edges(o, n, 1, None, N, N)
        with torch.no_grad():
            ref = layer(x_all.to(dev), ei)[sl].cpu()
        q.put((rank, rel(out, ref)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_host_pipelined_gcn_with_halo():
    """streaming.gcn_forward_host(partition=...) on 2 ranks (chunked copies, one halo exchange of the projected boundary
    rows, ghost-reading chunks deferred) == the owned rows of the monolithic layer output."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, e in res:
        assert e < 1e-5, res
