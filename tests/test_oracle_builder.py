"""Pins oracle/builder_oracle.py (the numpy restatement) to outputs of the UNMODIFIED reference
builder (tests/golden/*, produced by oracle/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import builder_oracle as bo


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def shipped(golden_dir):
    z = np.load(os.path.join(golden_dir, "shipped_mesh.npz"))
    mesh = dict(owner=z['owner'], neighbour=z['neighbour'], cell_centers=z['cell_centers'],
                internal_mask=z['internal_mask'], n_cells=int(z['n_cells']))
    gold = json.load(open(os.path.join(golden_dir, "builder_golden.json")))
    return mesh, gold


def test_loader_arrays_are_the_reference_ones(shipped):
    mesh, gold = shipped
    # SURVEY §8c goldens (reference loader output incl. its header-parsing quirk)
    assert sha(mesh['owner']) == "66dc8fdfc7796a9057cd15f706593b586b1171dcc52375f0d1a7533526bd0dd7"
    assert sha(mesh['neighbour']) == "da07d1efef539a8dfa73a775a31a1c6df4ad4d521d700f4ea85db6cae6b23ea6"
    assert mesh['n_cells'] == 49181 == gold['loader']['n_cells']
    assert mesh['owner'][:9].tolist() == [2, 0, 32, 64, 25012, 12225, 49180, 24170, 49180]


def _check(g, d):
    assert g['num_nodes'] == d['num_nodes']
    assert g['edge_index'].shape == (2, d['E'])
    assert g['edge_index'].dtype == np.int64
    assert sha(g['edge_index']) == d['edge_index_sha256']
    assert sha(g['edge_attr']) == d['edge_attr_sha256']
    assert sha(g['x']) == d['x_sha256']


def test_shipped_mode_A(shipped):
    mesh, gold = shipped
    g = bo.build_graph(mesh, node_features=mesh['cell_centers'], filter_internal=True, n_internal_cells=12225)
    _check(g, gold['mode_A'])
    assert gold['mode_A']['edge_index_sha256'] == "6ea6fc598a447f42e4802a9d16aa4a8feb70d3522a8b0413d7f10cee4737ae0d"
    assert sha(bo.compute_edge_attributes(mesh['cell_centers'], g['edge_index'])) == \
        gold['compute_edge_attributes_mode_A_sha256']


def test_shipped_mode_B(shipped):
    mesh, gold = shipped
    _check(bo.build_graph(mesh, filter_internal=True), gold['mode_B'])
    assert gold['mode_B']['edge_index_sha256'] == "cd2136f77642dbf5569d187dff5c3c8f4eb13435cf00a626b0fea93eeb0d85a4"


def test_shipped_mode_C(shipped):
    mesh, gold = shipped
    _check(bo.build_graph(mesh, node_features=mesh['cell_centers']), gold['mode_C'])
    assert gold['mode_C']['edge_index_sha256'] == "4baffe2f6535b17c2e43e9accd15621b9329cdc764dcfe38761302ce33280fce"
    assert gold['mode_C']['num_nodes'] == 49181 and gold['mode_C']['E'] == 110302


def test_shipped_build_edge_index(shipped):
    mesh, gold = shipped
    ei = bo.build_edge_index(mesh['owner'], mesh['neighbour'])
    assert list(ei.shape) == gold['build_edge_index']['shape']
    assert sha(ei) == gold['build_edge_index']['sha256']


def _toy_mesh(t, **over):
    m = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in t['mesh'].items()}
    m['owner'] = m['owner'].astype(np.int32)
    m['neighbour'] = m['neighbour'].astype(np.int32)
    m['cell_centers'] = np.asarray(m['cell_centers'], dtype=np.float64)
    m.update(over)
    return m


def _eq_full(g, d):
    assert g['num_nodes'] == d['num_nodes']
    assert g['edge_index'].tolist() == (d['edge_index'] if d['edge_index'] else [[], []])
    np.testing.assert_array_equal(g['edge_attr'], np.asarray(d['edge_attr'], dtype=np.float32).reshape(-1, 4))
    np.testing.assert_array_equal(g['x'], np.asarray(d['x'], dtype=np.float32))


def test_toy_known_answers(golden_dir):
    t = json.load(open(os.path.join(golden_dir, "toy_golden.json")))
    m = _toy_mesh(t)
    # literal vectors from SURVEY §8c
    assert t['build_edge_index'] == [[0, 1, 0, 2, 1, 3, 2, 3, 0, 0, 1, 1, 2, 2, 3, 3, 5],
                                     [1, 0, 2, 0, 3, 1, 3, 2, 0, 0, 1, 1, 2, 2, 3, 3, 5]]
    assert t['mode_A_n5']['edge_index'] == [[0, 1, 0, 2, 1, 3, 2, 3, 4], [1, 0, 2, 0, 3, 1, 3, 2, 4]]
    assert t['mode_A_n3']['edge_index'] == [[0, 1, 0, 2], [1, 0, 2, 0]]
    assert t['mode_B']['edge_index'] == [[0, 1, 1, 2, 3], [1, 0, 2, 1, 3]]
    assert t['mode_C']['edge_index'][0][-1] == 4
    assert bo.build_edge_index(m['owner'], m['neighbour']).tolist() == t['build_edge_index']
    _eq_full(bo.build_graph(m), t['mode_C'])
    _eq_full(bo.build_graph(m, filter_internal=True, n_internal_cells=5), t['mode_A_n5'])
    _eq_full(bo.build_graph(m, filter_internal=True, n_internal_cells=3), t['mode_A_n3'])
    _eq_full(bo.build_graph(m, filter_internal=True, n_internal_cells=1), t['mode_A_n1'])
    _eq_full(bo.build_graph(m, filter_internal=True), t['mode_nofilter_fallback'])
    mb = _toy_mesh(t, internal_mask=np.asarray(t['mode_B_mask'], dtype=bool))
    _eq_full(bo.build_graph(mb, filter_internal=True), t['mode_B'])
    ms = _toy_mesh(t, n_cells=4)
    ms['cell_centers'] = ms['cell_centers'][:4]
    _eq_full(bo.build_graph(ms), t['mode_C_ncells4'])
    me = dict(owner=np.array([1, 1, 3], dtype=np.int32), neighbour=np.array([], dtype=np.int32),
              cell_centers=m['cell_centers'][:5], n_cells=5)
    _eq_full(bo.build_graph(me), t['mode_C_no_internal'])
    _eq_full(bo.build_graph(me, filter_internal=True, n_internal_cells=3), t['mode_A_no_internal_n3'])
    fd = dict(U=np.arange(18, dtype=np.float64).reshape(6, 3), p=np.arange(6, dtype=np.float64),
              nut=np.arange(6, dtype=np.float64) * 2)
    np.testing.assert_array_equal(bo.build_graph(m, field_data=fd)['x'],
                                  np.asarray(t['mode_C_fields_x'], dtype=np.float32))


def test_random_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_golden.npz"))
    for c in range(5):
        m = dict(owner=z[f"c{c}_owner"], neighbour=z[f"c{c}_neighbour"], cell_centers=z[f"c{c}_cc"],
                 n_cells=int(z[f"c{c}_ncells"]), internal_mask=z[f"c{c}_mask"])
        np.testing.assert_array_equal(bo.build_edge_index(m['owner'], m['neighbour']), z[f"c{c}_bei"])
        for tag, kw in (("C", {}), ("A", dict(filter_internal=True, n_internal_cells=int(z[f"c{c}_nA"]))),
                        ("B", dict(filter_internal=True))):
            g = bo.build_graph(m, **kw)
            assert g['num_nodes'] == int(z[f"c{c}_{tag}_n"])
            np.testing.assert_array_equal(g['edge_index'], z[f"c{c}_{tag}_ei"])
            np.testing.assert_array_equal(g['edge_attr'], z[f"c{c}_{tag}_ea"])
            np.testing.assert_array_equal(g['x'], z[f"c{c}_{tag}_x"])


def test_csr_definition_is_stable_sort():
    rng = np.random.default_rng(0)
    ei = rng.integers(0, 20, size=(2, 200))
    rowptr, col, eid = bo.csr_by_target(ei, 20)
    import torch
    order = torch.sort(torch.from_numpy(ei[1]), stable=True).indices.numpy()
    np.testing.assert_array_equal(eid, order)
    np.testing.assert_array_equal(col, ei[0][order])
    assert rowptr[-1] == 200
