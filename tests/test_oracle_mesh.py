"""Pins oracle/mesh_oracle.py (numpy restatement of openfoam_loader.py:191-248) to what the UNMODIFIED reference loader
computed for the shipped case (tests/golden/shipped_mesh.npz + shipped_polymesh.npz, oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import mesh_oracle as mo


@pytest.fixture(scope="module")
def shipped(golden_dir):
    m = np.load(os.path.join(golden_dir, "shipped_mesh.npz"))
    pm = np.load(os.path.join(golden_dir, "shipped_polymesh.npz"))
    return m, pm


def test_cell_centers_match_the_reference_loader(shipped):
    m, pm = shipped
    cc = mo.get_cell_centers(pm['points'], m['owner'], m['neighbour'], pm['face_pts'], pm['face_off'])
    ref = m['cell_centers']
    assert cc.shape == ref.shape == (49181, 3) and cc.dtype == np.float64
    # the reference adds a cell's vertices in CPython set order, the oracle in ascending id: last-bit differences only
    assert np.abs(cc - ref).max() <= 1e-13
    assert mo.num_cells(m['owner'], m['neighbour']) == int(m['n_cells'])


def test_internal_mask_matches_the_reference_loader(shipped):
    m, _ = shipped
    mask = mo.get_internal_cells(m['owner'], m['neighbour'])
    assert mask.dtype == bool and np.array_equal(mask, m['internal_mask'])


def test_toy_known_answer():
    # two unit cubes sharing the face x = 1 (vertices 1,4,7,10 ... here: a 3 x 2 x 2 vertex grid, id = x + 3 (y + 2 z))
    pts = np.array([(x, y, z) for z in (0, 1) for y in (0, 1) for x in (0, 1, 2)], dtype=np.float64)
    vid = lambda x, y, z: x + 3 * (y + 2 * z)   # noqa: E731
    shared = [vid(1, 0, 0), vid(1, 1, 0), vid(1, 1, 1), vid(1, 0, 1)]
    left = [vid(0, 0, 0), vid(0, 1, 0), vid(0, 1, 1), vid(0, 0, 1)]
    right = [vid(2, 0, 0), vid(2, 1, 0), vid(2, 1, 1), vid(2, 0, 1)]
    faces = [shared, left, right, [vid(0, 0, 0), vid(1, 0, 0), vid(1, 0, 1)]]        # last one: a triangle (ragged)
    owner = np.array([0, 0, 1, 0], dtype=np.int32)
    neighbour = np.array([1], dtype=np.int32)
    fp, fo = mo.flatten_faces(faces)
    assert fo.tolist() == [0, 4, 8, 12, 15]
    cc = mo.get_cell_centers(pts, owner, neighbour, fp, fo)
    assert np.allclose(cc, [[0.5, 0.5, 0.5], [1.5, 0.5, 0.5]], atol=0, rtol=0)
    assert mo.get_internal_cells(owner, neighbour).tolist() == [True, True]
    # a cell id that no face names keeps a zero centre (openfoam_loader.py:221-223)
    cc2 = mo.get_cell_centers(pts, np.array([0, 0, 3, 0], dtype=np.int32), neighbour, fp, fo)
    assert cc2.shape == (4, 3) and not cc2[2].any() and np.allclose(cc2[3], [2, 0.5, 0.5])
