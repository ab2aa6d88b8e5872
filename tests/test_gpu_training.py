"""SURVEY §8f-3 / §8f-4 (csrc/train_glue.cu): device collation of Batch.from_data_list, the reference's weighted MSE loss and
clip + Adam as fused kernels — each against the torch formulation the reference executes (train.py:155-189,
normalization.py:177-236)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _samples():
    from gnn_bfs_rans_b200.data import Data
    g = torch.Generator().manual_seed(0)
    out = []
    for n, e in ((500, 1800), (3, 2), (7000, 26000), (1, 0)):
        out.append(Data(x=torch.randn(n, 3, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g),
                        edge_attr=torch.randn(e, 4, generator=g), y=torch.randn(n, 7, generator=g), num_nodes=n))
    return out


def test_device_batch_collation_equals_host_collation():
    from gnn_bfs_rans_b200 import _lib
    from gnn_bfs_rans_b200.data import Batch
    ds = _samples()
    host = Batch.from_data_list(ds)
    host.keys()                                           # materialise on the host (PyG semantics, tested on the CPU)
    _lib.launch_count_reset()
    dev = Batch.from_data_list(ds).to('cuda')
    assert _lib.launch_count() == 1                       # one finalize kernel; the copies go sample -> slice directly
    for k in ('x', 'y', 'edge_attr', 'edge_index', 'batch', 'ptr'):
        assert dev[k].is_cuda and torch.equal(dev[k].cpu(), host[k]), k
    assert dev.num_nodes == host.num_nodes and dev.num_graphs == 4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("prw", [0.0, 0.1])
def test_weighted_mse_loss_matches_the_reference_formula(dtype, prw):
    from gnn_bfs_rans_b200.training import WeightedMSELoss
    torch.manual_seed(0)
    crit = WeightedMSELoss()
    n = 12225
    pred = (torch.randn(n, 7, device='cuda') * 2 + 0.3).to(dtype)
    tgt = torch.randn(n, 7, device='cuda').to(dtype)
    p1 = pred.clone().requires_grad_(True)
    l1 = crit(p1, tgt, pressure_ref_weight=prw)
    (l1 * 1.7).backward()
    p2 = pred.double().requires_grad_(True)
    l2 = crit._reference_formula(p2, tgt.double(), prw)
    (l2 * 1.7).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(float(l1) - float(l2)) <= tol * abs(float(l2))
    assert float((p1.grad.double() - p2.grad).abs().max() / p2.grad.abs().max()) <= tol
    again = crit(pred, tgt, pressure_ref_weight=prw)
    assert float(again) == float(l1)                      # deterministic reduction
    assert float(WeightedMSELoss(use_fieldwise=False)(pred.float(), tgt.float())) > 0      # element-wise variant: torch path


def test_fused_clip_adam_follows_torch_clip_and_adam():
    from gnn_bfs_rans_b200.training import FusedClipAdam
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 7)).cuda()
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    oa = torch.optim.Adam(a.parameters(), lr=3e-3, weight_decay=1e-2)
    ob = FusedClipAdam(b.parameters(), lr=3e-3, weight_decay=1e-2, max_grad_norm=1.0)
    x = torch.randn(256, 16, device='cuda')
    y = torch.randn(256, 7, device='cuda') * 5
    for it in range(12):
        oa.zero_grad(set_to_none=True)
        la = (a(x) - y).square().mean()
        la.backward()
        na = torch.nn.utils.clip_grad_norm_(a.parameters(), 1.0)
        oa.step()
        ob.zero_grad()
        lb = (b(x) - y).square().mean()
        lb.backward()
        ob.step()
        assert abs(float(ob.grad_norm) - float(na)) <= 1e-4 * float(na), it
        assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(la)), it
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa - pb).abs().max()) <= 2e-6 * max(float(pa.abs().max()), 1.0)
    assert float(ob.state[0]) == 12.0


def test_fused_loss_and_optimizer_inside_a_captured_train_step():
    """The cfg2-shaped step (FlowGNN GCN on a small hex mesh) with the fused criterion and FusedClipAdam captured as one CUDA
    graph follows the same step run eagerly with the reference's torch pieces."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200 import ops
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200.training import FusedClipAdam, WeightedMSELoss
    nx, ny, nz = 14, 11, 9
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    gen = torch.Generator(device='cuda').manual_seed(0)
    xs = [torch.rand(N, 3, device='cuda', generator=gen) for _ in range(5)]
    ys = [torch.rand(N, 7, device='cuda', generator=gen) for _ in range(5)]
    crit = WeightedMSELoss()

    def make():
        torch.manual_seed(0)
        return FlowGNN(3, 128, 7, 3, "GCN", dropout=0.0, fused_glue=True).cuda().train()
    m_e, m_g = make(), make()
    o_e = torch.optim.Adam(m_e.parameters(), lr=1e-3, weight_decay=1e-5)
    o_g = FusedClipAdam(m_g.parameters(), lr=1e-3, weight_decay=1e-5, max_grad_norm=1.0)
    gs = b2g.graphs.GraphedTrainStep(m_g, o_g, lambda out, t: crit(out, t, pressure_ref_weight=0.1), xs[0], ys[0], ei, warmup=3)
    for _ in range(3):
        o_e.zero_grad(set_to_none=True)
        l = crit._reference_formula(m_e(xs[0], ei), ys[0], 0.1); l.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), 1.0); o_e.step()
    for k in range(1, 5):
        o_e.zero_grad(set_to_none=True)
        le = crit._reference_formula(m_e(xs[k], ei), ys[k], 0.1); le.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), 1.0); o_e.step()
        lg = gs.step(xs[k], ys[k])
        assert abs(float(lg) - float(le)) <= 2e-4 * max(abs(float(le)), 1e-6), (k, float(lg), float(le))
