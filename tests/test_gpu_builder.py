"""K0 parity (bit-exact): CUDA builder through the GraphConstructor drop-in vs the golden outputs of
the unmodified reference (tests/golden) and vs the oracle on seeded random face lists."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sha(a):
    if isinstance(a, torch.Tensor):
        a = a.cpu().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def b2g():
    import gnn_bfs_rans_b200 as m
    return m


@pytest.fixture(scope="module")
def shipped(golden_dir):
    z = np.load(os.path.join(golden_dir, "shipped_mesh.npz"))
    mesh = dict(owner=z['owner'], neighbour=z['neighbour'], cell_centers=z['cell_centers'],
                internal_mask=z['internal_mask'], n_cells=int(z['n_cells']))
    return mesh, json.load(open(os.path.join(golden_dir, "builder_golden.json")))


def _check(g, d):
    assert g.num_nodes == d['num_nodes']
    assert tuple(g.edge_index.shape) == (2, d['E'])
    assert g.edge_index.dtype == torch.int64 and g.edge_index.is_contiguous()
    assert not g.edge_index.is_cuda           # reference returns CPU tensors
    assert sha(g.edge_index) == d['edge_index_sha256']
    assert sha(g.edge_attr) == d['edge_attr_sha256']
    assert sha(g.x) == d['x_sha256']


def test_shipped_modes(b2g, shipped):
    mesh, gold = shipped
    gc = b2g.GraphConstructor(mesh)
    _check(gc.build_graph(node_features=mesh['cell_centers'], filter_internal=True, n_internal_cells=12225), gold['mode_A'])
    _check(gc.build_graph(filter_internal=True), gold['mode_B'])
    _check(gc.build_graph(node_features=mesh['cell_centers']), gold['mode_C'])
    ei = gc.build_edge_index()
    assert list(ei.shape) == gold['build_edge_index']['shape'] and sha(ei) == gold['build_edge_index']['sha256']
    ga = gc.build_graph(node_features=mesh['cell_centers'], filter_internal=True, n_internal_cells=12225)
    assert sha(gc.compute_edge_attributes(ga.edge_index)) == gold['compute_edge_attributes_mode_A_sha256']


def test_device_resident_output(b2g, shipped):
    mesh, gold = shipped
    g = b2g.GraphConstructor(mesh).build_graph(node_features=mesh['cell_centers'], device='cuda')
    assert g.edge_index.is_cuda and g.edge_attr.is_cuda and g.x.is_cuda
    assert sha(g.edge_index) == gold['mode_C']['edge_index_sha256']


def _toy_mesh(t, **over):
    m = dict(owner=np.asarray(t['mesh']['owner'], dtype=np.int32), neighbour=np.asarray(t['mesh']['neighbour'], dtype=np.int32),
             cell_centers=np.asarray(t['mesh']['cell_centers'], dtype=np.float64), n_cells=t['mesh']['n_cells'])
    m.update(over)
    return m


def _eq_full(g, d):
    assert g.num_nodes == d['num_nodes']
    assert g.edge_index.tolist() == (d['edge_index'] if d['edge_index'] else [[], []])
    np.testing.assert_array_equal(g.edge_attr.numpy(), np.asarray(d['edge_attr'], dtype=np.float32).reshape(-1, 4))
    np.testing.assert_array_equal(g.x.numpy(), np.asarray(d['x'], dtype=np.float32))


def test_toy_known_answers(b2g, golden_dir):
    t = json.load(open(os.path.join(golden_dir, "toy_golden.json")))
    m = _toy_mesh(t)
    GC = b2g.GraphConstructor
    assert GC(m).build_edge_index().tolist() == t['build_edge_index']
    _eq_full(GC(m).build_graph(), t['mode_C'])
    _eq_full(GC(m).build_graph(filter_internal=True, n_internal_cells=5), t['mode_A_n5'])
    _eq_full(GC(m).build_graph(filter_internal=True, n_internal_cells=3), t['mode_A_n3'])
    _eq_full(GC(m).build_graph(filter_internal=True, n_internal_cells=1), t['mode_A_n1'])
    _eq_full(GC(m).build_graph(filter_internal=True), t['mode_nofilter_fallback'])
    mb = _toy_mesh(t, internal_mask=np.asarray(t['mode_B_mask'], dtype=bool))
    _eq_full(GC(mb).build_graph(filter_internal=True), t['mode_B'])
    ms = _toy_mesh(t, n_cells=4)
    ms['cell_centers'] = ms['cell_centers'][:4]
    _eq_full(GC(ms).build_graph(), t['mode_C_ncells4'])
    me = dict(owner=np.array([1, 1, 3], dtype=np.int32), neighbour=np.array([], dtype=np.int32),
              cell_centers=m['cell_centers'][:5], n_cells=5)
    _eq_full(GC(me).build_graph(), t['mode_C_no_internal'])
    _eq_full(GC(me).build_graph(filter_internal=True, n_internal_cells=3), t['mode_A_no_internal_n3'])
    fd = dict(U=np.arange(18, dtype=np.float64).reshape(6, 3), p=np.arange(6, dtype=np.float64),
              nut=np.arange(6, dtype=np.float64) * 2)
    np.testing.assert_array_equal(GC(m).build_graph(field_data=fd).x.numpy(), np.asarray(t['mode_C_fields_x'], dtype=np.float32))
    with pytest.raises(IndexError):
        GC(m).build_graph(filter_internal=True, n_internal_cells=7)     # reference: IndexError at :129
    with pytest.raises(ValueError):
        GC(dict(m, boundaries={})).get_boundary_mask("inlet")


def test_random_golden(b2g, golden_dir):
    z = np.load(os.path.join(golden_dir, "random_golden.npz"))
    for c in range(5):
        m = dict(owner=z[f"c{c}_owner"], neighbour=z[f"c{c}_neighbour"], cell_centers=z[f"c{c}_cc"],
                 n_cells=int(z[f"c{c}_ncells"]), internal_mask=z[f"c{c}_mask"])
        gc = b2g.GraphConstructor(m)
        np.testing.assert_array_equal(gc.build_edge_index().numpy(), z[f"c{c}_bei"])
        for tag, kw in (("C", {}), ("A", dict(filter_internal=True, n_internal_cells=int(z[f"c{c}_nA"]))),
                        ("B", dict(filter_internal=True))):
            g = gc.build_graph(**kw)
            assert g.num_nodes == int(z[f"c{c}_{tag}_n"])
            np.testing.assert_array_equal(g.edge_index.numpy(), z[f"c{c}_{tag}_ei"])
            np.testing.assert_array_equal(g.edge_attr.numpy(), z[f"c{c}_{tag}_ea"])
            np.testing.assert_array_equal(g.x.numpy(), z[f"c{c}_{tag}_x"])


@pytest.mark.parametrize("n_cells,n_int,n_bnd", [(5000, 20000, 3000), (300000, 900000, 50000), (1, 0, 1), (10, 0, 0)])
def test_vs_oracle_larger(b2g, n_cells, n_int, n_bnd):
    """Sizes past one scan tile (4096) and past the tile-scan block (1024 tiles), ragged tails."""
    from oracle import builder_oracle as bo
    rng = np.random.default_rng(n_cells)
    m = dict(owner=rng.integers(0, n_cells, n_int + n_bnd).astype(np.int32),
             neighbour=rng.integers(0, n_cells, n_int).astype(np.int32),
             cell_centers=rng.standard_normal((n_cells, 3)), n_cells=n_cells,
             internal_mask=rng.random(n_cells) < 0.5)
    gc = b2g.GraphConstructor(m)
    for kw in ({}, dict(filter_internal=True, n_internal_cells=max(1, n_cells // 3)), dict(filter_internal=True)):
        g, r = gc.build_graph(**kw), bo.build_graph(m, **kw)
        assert g.num_nodes == r['num_nodes']
        np.testing.assert_array_equal(g.edge_index.numpy(), r['edge_index'])
        np.testing.assert_array_equal(g.edge_attr.numpy(), r['edge_attr'])
        np.testing.assert_array_equal(g.x.numpy(), r['x'])


def test_full_size_properties(b2g):
    """cfg4-sized hex mesh (250x200x200 = 10M cells): size-independent properties of the output."""
    from gnn_bfs_rans_b200.synthetic import hex_mesh_faces
    nx, ny, nz = 250, 200, 200
    owner, nei = hex_mesh_faces(nx, ny, nz, device='cuda')
    N = nx * ny * nz
    ei = b2g.ops.build_graph_edges(owner, nei, 1, None, N, N)
    assert ei.shape == (2, 59_720_000)
    assert bool((ei[0, 0::2] == ei[1, 1::2]).all()) and bool((ei[1, 0::2] == ei[0, 1::2]).all())  # interleaved pairs
    assert bool((ei[0, 0::2] < ei[1, 0::2]).all())                 # owner < neighbour kept in order
    deg = torch.bincount(ei[1], minlength=N)
    assert int(deg.min()) == 3 and int(deg.max()) == 6 and int(deg.sum()) == 59_720_000
    # idempotence / determinism
    assert torch.equal(ei, b2g.ops.build_graph_edges(owner, nei, 1, None, N, N))
