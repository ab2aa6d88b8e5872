"""K2/K3 parity: b2g_seg_sum (generic kernel and the 512-byte-row fast path of aggregate_rows.cu) against the
index_select + scatter_add_ definition PyG's MessagePassing.propagate executes, in fp64.

Covers what the reference's meshes and PyG's semantics can produce: empty rows (isolated nodes), rows longer than the
8-entry straight-line path and longer than one 32-entry index window, self term with coefficient 1 / != 1, per-row and
per-edge scales, bias + ReLU epilogue, the band / panel row order (every row computed exactly once), ragged tails."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _csr(N, deg_fn, seed, band=None):
    rng = np.random.default_rng(seed)
    deg = deg_fn(rng, N).astype(np.int64)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    nnz = int(rowptr[-1])
    rows = np.repeat(np.arange(N), deg)
    if band is None:
        col = rng.integers(0, N, size=nnz)
    else:                                   # banded: neighbours at +-1, +-band-ish offsets
        off = rng.choice(np.array([-band, -1, 0, 1, band]), size=nnz)
        col = np.clip(rows + off, 0, N - 1)
    return (torch.from_numpy(rowptr).int().cuda(), torch.from_numpy(col).int().cuda(),
            torch.from_numpy(rows).long().cuda())


def _ref(x, rows, col, N, rs, cs, self_coef, bias, relu):
    xd = x.double()
    msg = xd[col.long()]
    if cs is not None:
        msg = msg * cs.double()[col.long()][:, None]
    out = torch.zeros(N, x.shape[1], dtype=torch.float64, device=x.device).index_add_(0, rows, msg)
    if rs is not None:
        out = out * rs.double()[:, None]
    if self_coef != 0.0:
        out = out + self_coef * xd[:N]
    if bias is not None:
        out = out + bias.double()
    if relu:
        out = out.clamp_min(0)
    return out


def _mesh_deg(rng, N):                      # mesh-like: 4..7, a few boundary / isolated rows
    d = rng.integers(4, 8, size=N)
    d[rng.integers(0, N, size=max(N // 50, 1))] = 0
    return d


def _mixed_deg(rng, N):                     # includes rows > 8 (cold path) and > 32 / > 64 (several index windows)
    d = rng.integers(0, 10, size=N)
    d[rng.integers(0, N, size=max(N // 100, 1))] = rng.integers(9, 40)
    d[rng.integers(0, N, size=5)] = 77
    d[0] = 8
    d[1] = 9
    d[2] = 32
    d[3] = 33
    return d


CASES = [
    # name,            rs,    cs,    self, bias,  relu
    ("plain",          False, False, 0.0, False, False),
    ("gcn_fwd",        True,  False, 0.0, True,  False),
    ("gcn_fwd_nobias", True,  False, 0.0, False, False),
    ("gcn_bwd",        True,  True,  0.0, False, False),
    ("gin",            False, False, 1.0, False, False),
    ("gin_eps",        False, False, 1.25, False, False),
    ("weighted_self",  True,  True,  0.5, False, False),   # falls back to the generic kernel where no fast variant exists
    ("relu_bias",      True,  False, 0.0, True,  True),
]


@pytest.mark.parametrize("dtype,F", [(torch.bfloat16, 256), (torch.float32, 128), (torch.float32, 256), (torch.bfloat16, 512),
                                     (torch.float32, 64), (torch.bfloat16, 128)])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("deg", ["mesh", "mixed"])
def test_seg_sum_matches_scatter_add(dtype, F, case, deg):
    from gnn_bfs_rans_b200 import ops
    name, use_rs, use_cs, self_coef, use_bias, relu = case
    N = 3001                                                  # not a multiple of the chunk: ragged tail
    rowptr, col, rows = _csr(N, _mesh_deg if deg == "mesh" else _mixed_deg, seed=hash((F, name, deg)) % 1000)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(N, F, device='cuda', generator=g).to(dtype)
    rs = (torch.rand(N, device='cuda', generator=g) + 0.5) if use_rs else None
    cs = (torch.rand(N, device='cuda', generator=g) + 0.5) if use_cs else None
    bias = torch.randn(F, device='cuda', generator=g) if use_bias else None
    out = ops.seg_sum(x, rowptr, col, N, rs, cs, self_coef, None, bias, relu=relu)
    ref = _ref(x, rows, col, N, rs, cs, self_coef, bias, relu)
    tol = 1e-5 if dtype == torch.float32 else 2e-2            # BASELINE.json gates: 1e-5 rel (fp32), 2e-2 (bf16)
    err = (out.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    assert err <= tol, f"{name} {dtype} F={F} {deg}: rel err {err:.2e}"


@pytest.mark.parametrize("dtype,F", [(torch.bfloat16, 256), (torch.float32, 256)])
@pytest.mark.parametrize("chunk,panel", [(32, 8192), (8, 64), (64, 256), (16, 16)])
def test_band_order_is_a_permutation_of_the_rows(dtype, F, chunk, panel):
    """The panel order only changes WHEN a row is computed: results are bit-identical to the linear sweep, for band
    sizes that do and do not divide the row count, panels that do not divide the band, and chunks == panels."""
    from gnn_bfs_rans_b200 import ops, _lib
    lib = _lib.load()
    for N, band in ((20000, 1000), (20011, 1037), (70000, 33000), (5000, 2600)):
        rowptr, col, rows = _csr(N, _mesh_deg, seed=N, band=band)
        x = torch.randn(N, F, device='cuda').to(dtype)
        lin = ops.seg_sum(x, rowptr, col, N, None, None, 0.0, None, None, band=0)
        out = torch.full_like(x, float('nan'))
        ops.seg_sum(x, rowptr, col, N, None, None, 0.0, None, None, out=out, band=band, tune=(0, chunk, panel))
        assert torch.equal(out, lin), f"N={N} band={band}"
        ref = _ref(x, rows, col, N, None, None, 0.0, None, False)
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        assert (out.double() - ref).abs().max().item() / ref.abs().max().item() <= tol


def test_fast_path_equals_generic_kernel_bitwise():
    """Same per-row summation order in both kernels -> identical bits (deterministic aggregation)."""
    from gnn_bfs_rans_b200 import ops, _lib
    lib = _lib.load()
    N = 4099
    rowptr, col, rows = _csr(N, _mixed_deg, seed=7)
    for dtype in (torch.bfloat16, torch.float32):
        x = torch.randn(N, 256, device='cuda').to(dtype)
        rs = torch.rand(N, device='cuda') + 0.5
        fast = ops.seg_sum(x, rowptr, col, N, rs, None, 0.0, None, None)
        again = ops.seg_sum(x, rowptr, col, N, rs, None, 0.0, None, None)
        assert torch.equal(fast, again)                       # run-to-run deterministic
        gen = ops.seg_sum(x, rowptr, col, N, rs, None, 0.0, None, None, tune=(1, 0, 0))     # the generic kernel
        assert torch.equal(fast, gen)


def test_strided_views_and_errors():
    from gnn_bfs_rans_b200 import ops
    N = 2048
    rowptr, col, rows = _csr(N, _mesh_deg, seed=3)
    big = torch.randn(N, 3 * 256, device='cuda').bfloat16()
    x = big[:, 256:512]                                       # a column slice of a wider GEMM output (row stride 768)
    out = ops.seg_sum(x, rowptr, col, N, None, None, 1.0, None, None)
    ref = _ref(x, rows, col, N, None, None, 1.0, None, False)
    assert (out.double() - ref).abs().max().item() / ref.abs().max().item() <= 2e-2
    with pytest.raises(RuntimeError):
        ops.seg_sum(torch.randn(N, 256).bfloat16(), rowptr, col, N)      # host tensor: no CPU fallback


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows_per_chunk", [1024, 5000, 1 << 19])
def test_host_pipelined_gcn_equals_device_forward(dtype, rows_per_chunk):
    """streaming.gcn_forward_host (chunked copies overlapped with the kernels) == layer(x.cuda(), ei.cuda()).cpu(),
    bit for bit, on a banded mesh and on a graph with no band structure (every chunk then waits for all of x)."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200 import ops
    torch.manual_seed(0)
    layer = b2g.nn.GCNConv(256, 256).cuda().to(dtype).eval()
    with torch.no_grad():
        layer.bias.uniform_(-1, 1)
    nx, ny, nz = 40, 30, 20
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei_mesh = ops.build_graph_edges(o, n, 1, None, N, N)
    ei_rand = torch.randint(0, N, (2, 6 * N), device='cuda')
    for ei in (ei_mesh, ei_rand):
        x = torch.randn(N, 256).to(dtype)
        with torch.no_grad():
            ref = layer(x.cuda(), ei).cpu()
        out = b2g.streaming.gcn_forward_host(layer, x, ei.cpu(), rows_per_chunk=rows_per_chunk)
        assert out.is_pinned() and torch.equal(out, ref)


@pytest.mark.parametrize("kind", ["GAT", "GIN", "Transformer"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_host_pipeline_all_layer_types(kind, dtype):
    """streaming.forward_host for GATConv / GINConv / TransformerConv: the row-chunked host pipeline (or, for configurations
    without one — fp32 GATConv — the whole-graph call between pinned copies) against layer(x.cuda(), ei.cuda()).cpu().  The
    chunked kernels do the same arithmetic per row; bf16 GAT goes through the fused kernel in both (bit-equal), the others
    may differ in the last bit where a GEMM tile boundary moves (tolerance 1e-6 fp32 / 1 bf16 ulp)."""
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    from gnn_bfs_rans_b200 import ops
    torch.manual_seed(0)
    F = 256
    layer = {"GAT": lambda: b2g.nn.GATConv(F, F, heads=4, concat=False),
             "GIN": lambda: b2g.nn.GINConv(torch.nn.Sequential(torch.nn.Linear(F, F), torch.nn.ReLU(), torch.nn.Linear(F, F))),
             "Transformer": lambda: b2g.nn.TransformerConv(F, F, heads=4, concat=False)}[kind]().cuda().to(dtype).eval()
    nx, ny, nz = 40, 30, 20
    N = nx * ny * nz
    o, n = hex_mesh_faces(nx, ny, nz, device='cuda')
    ei_mesh = ops.build_graph_edges(o, n, 1, None, N, N)
    ei_rand = torch.cat([torch.randint(0, N, (2, 5 * N), device='cuda'), torch.stack([torch.randint(0, N, (50,), device='cuda'),
                                                                                     torch.full((50,), 7, device='cuda')])], 1)
    for ei in (ei_mesh, ei_rand):
        x = torch.randn(N, F).to(dtype)
        with torch.no_grad():
            ref = layer(x.cuda(), ei).cpu()
        for rpc in (2048, 1 << 19):
            out = b2g.streaming.forward_host(layer, x, ei.cpu(), rows_per_chunk=rpc)
            assert out.is_pinned() and out.shape == ref.shape
            err = float((out.double() - ref.double()).abs().max() / ref.double().abs().max())
            assert err <= (1e-6 if dtype == torch.float32 else 8e-3), (kind, rpc, err)
    with pytest.raises(RuntimeError):
        b2g.streaming.forward_host(layer.train(), x, ei.cpu())
