"""The hot path behind its caller: FlowGNN (mirror of gnn_model.py:14-220) forward / gradients vs the
fp64 oracle for every layer type, the drop-in module registration, and a short training run."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    b = b.double().cpu()
    return float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def grid_graph(nx, ny):
    ids = np.arange(nx * ny).reshape(ny, nx)
    o = np.concatenate([ids[:, :-1].ravel(), ids[:-1, :].ravel()])
    n = np.concatenate([ids[:, 1:].ravel(), ids[1:, :].ravel()])
    order = np.lexsort((n, o))
    return o[order].astype(np.int32), n[order].astype(np.int32)


@pytest.mark.parametrize("layer_type", ["GCN", "GAT", "GIN", "Transformer", "Transformer+edge", "GAT+edge"])
@pytest.mark.parametrize("training", [False, True])
def test_flowgnn_matches_oracle(layer_type, training):
    # "Transformer+edge": FlowGNN(edge_dim=4), the Transformer layers consume the edge_attr the reference passes (§8f-2)
    edge_dim = 4 if layer_type.endswith("+edge") else None
    layer_type = layer_type.split("+")[0]
    hidden = 128 if edge_dim else 64
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200 import GraphConstructor
    from oracle import layers_oracle as lo
    nx, ny = 40, 30
    o, n = grid_graph(nx, ny)
    cc = np.random.default_rng(0).standard_normal((nx * ny, 3))
    g = GraphConstructor(dict(owner=o, neighbour=n, cell_centers=cc, n_cells=nx * ny)).build_graph(
        filter_internal=True, n_internal_cells=nx * ny)
    torch.manual_seed(0)
    model = FlowGNN(3, hidden, 7, 3, layer_type, dropout=0.0, edge_dim=edge_dim).cuda()
    model.train(training)
    x = g.x.cuda().requires_grad_(True)
    out = model(x, g.edge_index.cuda(), g.edge_attr.cuda())
    out.square().mean().backward()
    pnames = {k for k, _ in model.named_parameters()}
    p = {k: v.detach().double().cpu().requires_grad_(k in pnames) for k, v in model.state_dict().items()}
    x64 = g.x.double().requires_grad_(True)
    ref = lo.flow_gnn_forward(x64, g.edge_index, p, layer_type, training=training,
                              edge_attr=g.edge_attr.double() if edge_dim else None)
    ref.square().mean().backward()
    if edge_dim:
        assert model.gnn_layers[0].lin_edge.weight.grad is not None and 'gnn_layers.0.lin_edge.weight' in p
        assert float(model.gnn_layers[0].lin_edge.weight.grad.abs().max()) > 0
    # whole-model gate: L layers + BatchNorm + head in fp32 (torch ops of the caller included), so the
    # per-layer 1e-5 compounds; gradients are gauged against the largest gradient of the model because
    # several (biases in front of a BatchNorm) are zero in exact arithmetic.
    assert rel(out.detach(), ref.detach()) < 2e-5
    gx, gr = x.grad.double().cpu(), x64.grad
    # A ReLU whose fp32 pre-activation lands within rounding of 0 flips against fp64 and moves the gradient
    # rows of that node's L-hop neighbourhood (~13-25 nodes of this grid for one flip) by O(1) whatever the
    # kernels do.  Gate: at most 3% of the rows may deviate by more than 1e-4 of the largest gradient, and
    # all other rows must agree to 1e-4 in L2.
    row_err = (gx - gr).abs().max(1).values
    bad = row_err > 1e-4 * gr.abs().max()
    assert int(bad.sum()) <= max(3, (3 * gx.shape[0]) // 100), int(bad.sum())
    assert float((gx - gr)[~bad].norm() / gr[~bad].norm()) < 1e-4
    scale = max(float(p[n].grad.abs().max()) for n in pnames if p[n].grad is not None)
    for name, par in model.named_parameters():
        if par.grad is not None and p[name].grad is not None:
            err = float((par.grad.double().cpu() - p[name].grad).abs().max())
            # 5e-3: a single flipped ReLU (see above) enters every parameter sum of the layers below it
            assert err / max(float(p[name].grad.abs().max()), 1e-4 * scale) < 5e-3, name


def test_dropin_registers_reference_import_names():
    import importlib
    import sys
    from gnn_bfs_rans_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data", "graph_constructor")}
    try:
        dropin.install()
        nn = importlib.import_module("torch_geometric.nn")
        from torch_geometric.nn import MessagePassing, global_mean_pool, GCNConv, GATConv, GINConv, TransformerConv, BatchNorm  # noqa
        from torch_geometric.data import Data, Batch  # noqa
        from graph_constructor import GraphConstructor  # noqa
        assert nn.GCNConv.__module__.endswith("nn")
        d1 = Data(x=torch.randn(3, 2), edge_index=torch.tensor([[0, 1], [1, 2]]), num_nodes=3)
        d2 = Data(x=torch.randn(2, 2), edge_index=torch.tensor([[0], [1]]), num_nodes=2)
        b = Batch.from_data_list([d1, d2]).to('cuda')
        assert b.edge_index.tolist() == [[0, 1, 3], [1, 2, 4]] and b.batch.tolist() == [0, 0, 0, 1, 1]
        assert b.x.is_cuda and b.num_nodes == 5
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_short_training_run_reduces_loss():
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    o, n = grid_graph(50, 40)
    N = 2000
    ei = torch.cat([torch.tensor(np.stack([o, n])), torch.tensor(np.stack([n, o]))], 1).long().cuda()
    torch.manual_seed(0)
    x = torch.rand(N, 3, device='cuda')
    y = torch.stack([torch.sin(3 * x[:, 0]), torch.cos(2 * x[:, 1]), x[:, 0] * x[:, 1], x[:, 2], x.sum(1), x[:, 0] ** 2, x[:, 1]], 1)
    for lt in ("GCN", "GAT"):
        model = FlowGNN(3, 64, 7, 3, lt, dropout=0.1).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=3e-3)
        losses = []
        for _ in range(60):
            opt.zero_grad()
            loss = (model(x, ei) - y).square().mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            losses.append(float(loss))
        assert losses[-1] < 0.5 * losses[0], (lt, losses[0], losses[-1])


@pytest.mark.parametrize("layer_type", ["GCN", "GAT", "GIN", "Transformer"])
def test_cuda_graph_forward_equals_eager(layer_type):
    """graphs.GraphedForward replays the eval forward of the static-mesh model from one CUDA graph: bit-identical to the
    eager call, also for new inputs copied into the captured buffer."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200 import ops
    nx, ny, nz = 16, 12, 10
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    torch.manual_seed(0)
    model = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.1).cuda().eval()
    x0 = torch.rand(N, 3, device='cuda')
    gf = b2g.graphs.GraphedForward(model, x0, ei)
    for seed in (1, 2):
        x = torch.rand(N, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(seed))
        with torch.no_grad():
            ref = model(x, ei)
        assert torch.equal(gf(x), ref)
    with pytest.raises(ValueError):
        gf(torch.rand(N + 1, 3, device='cuda'))
    with pytest.raises(RuntimeError):
        b2g.graphs.GraphedForward(model.train(), x0, ei)


@pytest.mark.parametrize("layer_type", ["GCN", "GAT"])
def test_cuda_graph_train_step(layer_type):
    """graphs.GraphedTrainStep: (i) with dropout 0 the captured step follows the eager step (same kernels; Adam in
    capturable mode) — loss trajectories agree; (ii) with dropout > 0 and a zero learning rate two replays on the same
    sample give different losses (the device-side dropout epoch advances inside the graph), while eager-mode seeds are
    untouched by the epoch mechanism until it is advanced."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200 import ops
    nx, ny, nz = 14, 11, 9
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    gen = torch.Generator(device='cuda').manual_seed(0)
    xs = [torch.rand(N, 3, device='cuda', generator=gen) for _ in range(6)]
    ys = [torch.rand(N, 7, device='cuda', generator=gen) for _ in range(6)]
    loss_fn = lambda out, y: (out - y).square().mean()

    def make(p):
        torch.manual_seed(0)
        m = FlowGNN(3, 128, 7, 3, layer_type, dropout=p, fused_glue=True).cuda().train()
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)

    # (i) trajectories
    m_e, o_e = make(0.0)
    m_g, o_g = make(0.0)
    m_g.load_state_dict(m_e.state_dict())
    gs = b2g.graphs.GraphedTrainStep(m_g, o_g, loss_fn, xs[0], ys[0], ei, max_grad_norm=1.0, warmup=3)
    for _ in range(3):                                   # the 3 warm-up steps on sample 0 (capturing executes nothing)
        o_e.zero_grad(set_to_none=True)
        l = loss_fn(m_e(xs[0], ei), ys[0]); l.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), 1.0); o_e.step()
    for k in range(1, 6):
        o_e.zero_grad(set_to_none=True)
        le = loss_fn(m_e(xs[k], ei), ys[k]); le.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), 1.0); o_e.step()
        lg = gs.step(xs[k], ys[k])
        assert abs(float(lg) - float(le)) <= 1e-4 * max(abs(float(le)), 1e-6), (k, float(lg), float(le))
    # (ii) fresh dropout masks per replay
    torch.manual_seed(1)
    m_d = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.3, fused_glue=True).cuda().train()
    o_d = torch.optim.Adam(m_d.parameters(), lr=0.0, capturable=True)
    gd = b2g.graphs.GraphedTrainStep(m_d, o_d, loss_fn, xs[0], ys[0], ei, warmup=2)
    l1 = float(gd.step(xs[0], ys[0])); l2 = float(gd.step(xs[0], ys[0]))
    assert l1 != l2 and abs(l1 - l2) < 0.5 * abs(l1)


@pytest.mark.parametrize("layer_type", ["GCN", "GAT"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [False, True])
def test_flowgnn_on_the_shipped_graph_cfg1_cfg2(layer_type, dtype, training, golden_dir):
    """BASELINE cfg1 / cfg2 as configured: the shipped BFS case's train graph (12 225 cells, mode A), FlowGNN hidden 128,
    L = 4, GCN and GAT (4 heads), eval and train mode (batch statistics), fp32 and bf16, against oracle.flow_gnn_forward in
    fp64 on identical weights and inputs.
    fp32: forward <= 2e-5 (four layers + BatchNorm + head compound the per-layer 1e-5); input gradients: rows touched by a
    ReLU that flips between fp32 and fp64 are counted (<= 5 %), every other row agrees to 1e-4 (eval) / 3e-4 (train) in relative L2.
    bf16: forward <= 2e-2, or — train mode, where BatchNorm divides by the batch standard deviation of channels that the
    rank-3 input leaves almost constant — <= 1.5 x the error the SAME oracle makes when executed in bf16 on the CPU (measured:
    8e-2 GCN / 2.2e-1 GAT for the CPU bf16 oracle, 7e-2 / 2.2e-1 for the kernels)."""
    import os
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200 import GraphConstructor
    from oracle import layers_oracle as lo
    z = np.load(os.path.join(golden_dir, "shipped_mesh.npz"))
    mesh = dict(owner=z['owner'], neighbour=z['neighbour'], cell_centers=z['cell_centers'], n_cells=int(z['n_cells']))
    g = GraphConstructor(mesh).build_graph(node_features=mesh['cell_centers'], filter_internal=True, n_internal_cells=12225)
    assert g.num_nodes == 12225 and g.edge_index.shape[1] == 48330
    torch.manual_seed(0)
    model = FlowGNN(3, 128, 7, 4, layer_type, dropout=0.0).cuda().to(dtype)
    model.train(training)
    x = g.x.to(dtype)
    xg = x.cuda().requires_grad_(True)
    out = model(xg, g.edge_index.cuda(), g.edge_attr.cuda())
    out.float().square().mean().backward()
    pn = {k for k, _ in model.named_parameters()}
    sd = model.state_dict()
    p = {k: v.detach().double().cpu().requires_grad_(k in pn) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    ref = lo.flow_gnn_forward(x64, g.edge_index, p, layer_type, training=training)
    ref.square().mean().backward()
    e_fwd = rel(out.detach(), ref.detach())
    if dtype == torch.float32:
        assert e_fwd < 2e-5, e_fwd
        gx, gr = xg.grad.double().cpu(), x64.grad
        bad = (gx - gr).abs().max(1).values > 1e-4 * gr.abs().max()
        assert int(bad.sum()) <= (5 * gx.shape[0]) // 100, int(bad.sum())
        # train mode: 1 / sqrt(batch variance) of the almost-constant channels amplifies fp32 rounding (forward 4e-6 -> 1e-4 here)
        assert float((gx - gr)[~bad].norm() / gr[~bad].norm()) < (3e-4 if training else 1e-4)
        scale = max(float(p[n].grad.abs().max()) for n in pn if p[n].grad is not None)
        for name, par in model.named_parameters():
            if par.grad is not None and p[name].grad is not None:
                err = float((par.grad.double().cpu() - p[name].grad).abs().max())
                # gauge: 1e-3 of the model's largest gradient for the tensors whose exact gradient is zero (biases in front
                # of a train-mode BatchNorm); a flipped ReLU enters every parameter sum of the layers below it
                assert err / max(float(p[name].grad.abs().max()), 1e-3 * scale) < 1.5e-2, (name, err)
    else:
        bound = 2e-2
        if training:
            with torch.no_grad():
                pb = {k: v.detach().cpu() for k, v in sd.items()}
                ob = lo.flow_gnn_forward(x, g.edge_index, pb, layer_type, training=True)
            bound = max(bound, 1.5 * rel(ob, ref.detach()))
        assert e_fwd < bound, (e_fwd, bound)
        assert bool(torch.isfinite(xg.grad.float()).all())
