"""BatchNorm (+ fused residual / ReLU / dropout) kernels of csrc/bn.cu against torch.nn.BatchNorm1d semantics in fp64
(torch_geometric.nn.BatchNorm wraps BatchNorm1d: gnn_model.py:87,188; the glue is gnn_model.py:184-192)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def _rel(a, b):
    b = b.double()
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,C", [(1, 64), (7, 8), (1000, 128), (4099, 256), (300, 1024)])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_matches_torch_fp64(dtype, N, C, training):
    import gnn_bfs_rans_b200 as b2g
    torch.manual_seed(N + C)
    bn = b2g.nn.BatchNorm(C).cuda()
    ref = torch.nn.BatchNorm1d(C).cuda().double()
    with torch.no_grad():
        bn.module.weight.uniform_(0.5, 1.5); bn.module.bias.uniform_(-0.5, 0.5)
        bn.module.running_mean.uniform_(-1, 1); bn.module.running_var.uniform_(0.5, 2)
        ref.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in bn.module.state_dict().items()})
    bn = bn.to(dtype)
    bn.train(training); ref.train(training)
    if training and N == 1:
        pytest.skip("BatchNorm1d rejects a single row in training mode")
    x = (torch.randn(N, C, device='cuda') * 2 + 3).to(dtype).requires_grad_(True)     # |mean| > std: shifted sums matter
    xr = x.detach().double().requires_grad_(True)
    y = bn(x)
    yr = ref(xr)
    assert y.dtype == dtype and _rel(y, yr) <= TOL[dtype]
    g = torch.randn(N, C, device='cuda').to(dtype)
    y.backward(g)
    yr.backward(g.double())
    gtol = TOL[dtype] * (1 if dtype == torch.float32 else 2)
    assert _rel(x.grad, xr.grad) <= gtol
    assert _rel(bn.module.weight.grad, ref.weight.grad) <= gtol and _rel(bn.module.bias.grad, ref.bias.grad) <= gtol
    if training:                                                                      # running statistics, momentum 0.1, unbiased var
        assert _rel(bn.module.running_mean, ref.running_mean) <= TOL[dtype]
        assert _rel(bn.module.running_var, ref.running_var) <= TOL[dtype]
        assert int(bn.module.num_batches_tracked) == int(ref.num_batches_tracked)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_residual_bn_relu(dtype):
    """relu(BN(h + h_new)) fused == the four torch ops of gnn_model.py:184-190 (p = 0), outputs and all gradients."""
    from gnn_bfs_rans_b200 import functional as Fn
    N, C = 3001, 256
    torch.manual_seed(0)
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    ref = torch.nn.BatchNorm1d(C).cuda().double()
    ref.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in bn.state_dict().items()})
    bn = bn.to(dtype)
    h = torch.randn(N, C, device='cuda').to(dtype).requires_grad_(True)
    hn = torch.randn(N, C, device='cuda').to(dtype).requires_grad_(True)
    y = Fn.batch_norm(h, hn, bn, relu=True, p_drop=0.0)
    s = (h.detach().float() + hn.detach().float()).to(dtype).double().requires_grad_(True)   # s is stored in `dtype`
    yr = torch.relu(ref(s))
    assert _rel(y, yr) <= TOL[dtype]
    g = torch.randn(N, C, device='cuda').to(dtype)
    y.backward(g)
    yr.backward(g.double())
    gtol = TOL[dtype] * (1 if dtype == torch.float32 else 2)
    if dtype == torch.float32:
        assert _rel(h.grad, s.grad) <= gtol
    else:   # bf16 rounds pre-activations near 0 across the ReLU kink: isolated entries flip, so gate the L2 error
        assert float((h.grad.double() - s.grad).norm() / s.grad.norm()) <= gtol
    assert torch.equal(h.grad, hn.grad)
    assert _rel(bn.weight.grad, ref.weight.grad) <= gtol and _rel(bn.bias.grad, ref.bias.grad) <= gtol


def test_fused_dropout_statistics_and_backward_mask():
    from gnn_bfs_rans_b200 import functional as Fn
    N, C, p = 20000, 256, 0.25
    bn = torch.nn.BatchNorm1d(C).cuda().train()
    h = torch.randn(N, C, device='cuda').requires_grad_(True)
    hn = torch.randn(N, C, device='cuda')
    y = Fn.batch_norm(h, hn, bn, relu=True, p_drop=p)
    y0 = Fn.batch_norm(h.detach(), hn, bn, relu=True, p_drop=0.0)
    active = y0 > 0
    kept = (y > 0) & active
    frac = kept.sum().item() / active.sum().item()
    assert abs(frac - (1 - p)) < 5e-3                                   # keep probability
    assert torch.allclose(y[kept], y0[kept] / (1 - p), rtol=1e-5, atol=1e-6)   # inverted-dropout scaling
    assert (y[~kept] == 0).all()
    y.backward(torch.ones_like(y))
    assert h.grad is not None and torch.isfinite(h.grad).all()
    y2 = Fn.batch_norm(h.detach(), hn, bn, relu=True, p_drop=p)
    assert not torch.equal(y2, y.detach())                              # a fresh mask per call


@pytest.mark.parametrize("layer_type", ["GCN", "GAT"])
def test_flowgnn_fused_glue_equals_unfused(layer_type):
    """FlowGNN(fused_glue=True) == FlowGNN (eval and train with p = 0): same model, fewer passes."""
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200 import ops
    nx, ny, nz = 12, 10, 9
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei = ops.build_graph_edges(o, n, 1, None, N, N)
    torch.manual_seed(0)
    a = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.0).cuda().train()
    b = FlowGNN(3, 128, 7, 3, layer_type, dropout=0.0, fused_glue=True).cuda().train()
    b.load_state_dict(a.state_dict())
    x = torch.rand(N, 3, device='cuda')
    ya, yb = a(x, ei), b(x, ei)
    assert _rel(yb, ya) <= 2e-5
    ya.square().mean().backward(); yb.square().mean().backward()
    for (na, pa), (nb_, pb) in zip(a.named_parameters(), b.named_parameters()):
        scale = max(float(pa.grad.abs().max()), 1e-12)
        assert float((pa.grad - pb.grad).abs().max()) / scale <= 5e-4, na
    for (na, ba), (nb_, bb) in zip(a.named_buffers(), b.named_buffers()):
        if ba.dtype.is_floating_point:
            assert _rel(bb, ba) <= 1e-5, na
