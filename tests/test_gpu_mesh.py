"""Device mesh ingest (csrc/mesh.cu through the C ABI) against oracle/mesh_oracle.py and the reference loader's own
arrays for the shipped case.  Gate: internal mask / n_cells bit-exact; cell centres 1e-13 absolute (fp64; the summation
order of the reference is CPython's set order, see the oracle header)."""
import os

import numpy as np
import pytest
import torch

from oracle import mesh_oracle as mo

pytestmark = pytest.mark.gpu


def _b2g():
    import gnn_bfs_rans_b200 as b2g
    return b2g


def test_shipped_case_matches_reference_loader(golden_dir):
    b2g = _b2g()
    m = np.load(os.path.join(golden_dir, "shipped_mesh.npz"))
    pm = np.load(os.path.join(golden_dir, "shipped_polymesh.npz"))
    faces = (pm['face_pts'], pm['face_off'])
    d = b2g.mesh.derive_mesh(pm['points'], m['owner'], m['neighbour'], faces)
    assert d['n_cells'] == int(m['n_cells']) == 49181
    assert isinstance(d['cell_centers'], np.ndarray) and d['cell_centers'].dtype == np.float64
    assert np.abs(d['cell_centers'] - m['cell_centers']).max() <= 1e-13          # the unmodified reference's output
    assert np.array_equal(d['internal_mask'], m['internal_mask'])
    assert d['n_internal_cells'] == int(m['internal_mask'].sum())
    # bit-equal with the oracle (same ascending-vertex summation order)
    cc_o = mo.get_cell_centers(pm['points'], m['owner'], m['neighbour'], pm['face_pts'], pm['face_off'])
    assert np.array_equal(d['cell_centers'], cc_o)
    # the derived dict drops into the graph builder like loader.load_mesh()'s
    mesh = dict(owner=m['owner'], neighbour=m['neighbour'], **d)
    g = b2g.GraphConstructor(mesh).build_graph(node_features=mesh['cell_centers'], filter_internal=True,
                                               n_internal_cells=12225)
    assert g.edge_index.shape == (2, 48330)


def _random_polymesh(seed, n_cells, n_points, n_faces, n_int, max_len):
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(n_points, 3))
    faces = [rng.integers(0, n_points, size=rng.integers(3, max_len + 1)).tolist() for _ in range(n_faces)]
    owner = rng.integers(0, n_cells, size=n_faces).astype(np.int32)
    neighbour = rng.integers(0, n_cells, size=n_int).astype(np.int32)
    return pts, owner, neighbour, faces


@pytest.mark.parametrize("seed,n_cells,n_points,n_faces,n_int,max_len", [
    (0, 50, 40, 300, 120, 6),          # heavy vertex duplication inside cells
    (1, 5000, 20000, 30000, 14000, 5),
    (2, 300, 1000, 200, 1, 8),         # one internal face; many cells without any face (zero centres)
    (3, 7, 9, 400, 400, 12),           # long per-cell lists (hundreds of slots)
])
def test_random_ragged_polymesh(seed, n_cells, n_points, n_faces, n_int, max_len):
    b2g = _b2g()
    pts, owner, neighbour, faces = _random_polymesh(seed, n_cells, n_points, n_faces, n_int, max_len)
    fp, fo = mo.flatten_faces(faces)
    cc_o = mo.get_cell_centers(pts, owner, neighbour, fp, fo)
    cc = b2g.mesh.get_cell_centers(pts, owner, neighbour, faces)        # ragged list form
    cc2 = b2g.mesh.get_cell_centers(torch.from_numpy(pts).cuda(), owner, neighbour, (fp, fo))
    assert torch.equal(cc, cc2)
    assert cc.shape == cc_o.shape
    assert np.array_equal(cc.cpu().numpy(), cc_o)                        # same summation order as the oracle: bit-equal
    mask = b2g.mesh.get_internal_cells(owner, neighbour).cpu().numpy()
    assert np.array_equal(mask, mo.get_internal_cells(owner, neighbour))
    # independent check in plain Python sets, the way the reference collects a cell's vertices (:203-214)
    for c in range(0, cc_o.shape[0], max(1, cc_o.shape[0] // 25)):
        ids = {p for f, o in zip(faces, owner) if o == c for p in f} | \
              {p for f, o in zip(faces[:n_int], neighbour) if o == c for p in f}
        want = np.mean(pts[sorted(ids)], axis=0) if ids else np.zeros(3)
        assert np.abs(cc_o[c] - want).max() <= 1e-13


def test_errors_follow_the_reference():
    b2g = _b2g()
    pts = np.zeros((4, 3))
    faces = [[0, 1, 2], [1, 2, 3]]
    with pytest.raises(IndexError):     # owner names a face that `faces` does not hold (faces[i], :205)
        b2g.mesh.get_cell_centers(pts, np.array([0, 0, 0], dtype=np.int32), np.array([0], dtype=np.int32), faces)
    with pytest.raises(IndexError):     # vertex id past the points array (points[idx], :219)
        b2g.mesh.get_cell_centers(pts, np.array([0, 1], dtype=np.int32), np.array([1], dtype=np.int32), [[0, 1, 9], [1, 2, 3]])
    with pytest.raises(IndexError):     # neighbour longer than owner (owner[i], :244)
        b2g.mesh.get_internal_cells(np.array([0], dtype=np.int32), np.array([0, 0], dtype=np.int32))
    with pytest.raises(ValueError):     # np.max of an empty neighbour list (:197)
        b2g.mesh.get_cell_centers(pts, np.array([0, 0], dtype=np.int32), np.zeros(0, dtype=np.int32), faces)
    with pytest.raises(RuntimeError):   # no CPU fallback
        b2g.ops.mesh_num_cells(torch.zeros(3, dtype=torch.int32), torch.zeros(1, dtype=torch.int32))


def test_hex_block_known_answer():
    """Structured block: the centre of cell (ix, iy, iz) is exactly (ix + .5, iy + .5, iz + .5) in fp64."""
    b2g = _b2g()
    from gnn_bfs_rans_b200.synthetic import hex_cell_centers, hex_polymesh
    nx, ny, nz = 64, 48, 40
    pts, own, nbr, fp, fo = hex_polymesh(nx, ny, nz, "cuda")
    d = b2g.mesh.derive_mesh(pts, own, nbr, (fp, fo), as_numpy=False)
    assert d["n_cells"] == nx * ny * nz and d["n_internal_cells"] == nx * ny * nz
    assert np.array_equal(d["cell_centers"].cpu().numpy(), hex_cell_centers(nx, ny, nz))
