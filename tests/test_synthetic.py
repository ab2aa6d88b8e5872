"""Synthetic mesh generators used by bench.py / the GPU tests (CPU only)."""
import numpy as np

from gnn_bfs_rans_b200.synthetic import (delaunay_dual_faces, hex_cell_centers, hex_mesh_faces, hex_polymesh,
                                         hilbert_index_2d, hilbert_renumber_2d)
from oracle import mesh_oracle as mo


def test_hilbert_index_is_a_space_filling_bijection():
    b = 5
    n = 1 << b
    X, Y = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    d = hilbert_index_2d(X.ravel(), Y.ravel(), b)
    assert sorted(d.tolist()) == list(range(n * n))
    o = np.argsort(d)
    px, py = X.ravel()[o], Y.ravel()[o]
    assert (np.abs(np.diff(px)) + np.abs(np.diff(py)) == 1).all()          # consecutive cells share an edge


def test_hilbert_renumbering_preserves_the_mesh():
    own, nbr, cen = delaunay_dual_faces(3000, seed=1)
    o2, n2, c2, order = hilbert_renumber_2d(own, nbr, cen)
    assert (o2 < n2).all() and (np.diff(o2.astype(np.int64)) >= 0).all()
    assert np.array_equal(c2, cen[order])
    old = {(min(a, b), max(a, b)) for a, b in zip(order[o2].tolist(), order[n2].tolist())}
    assert old == {(int(a), int(b)) for a, b in zip(own, nbr)}               # the same faces, renamed
    assert np.abs(o2.astype(np.int64) - n2).mean() < np.abs(own.astype(np.int64) - nbr).mean()


def test_hex_polymesh_is_consistent_with_the_face_lists():
    p, o, n, fp, fo = hex_polymesh(5, 4, 3)
    o2, n2 = hex_mesh_faces(5, 4, 3, boundary=True)
    assert np.array_equal(o.numpy(), o2.numpy()) and np.array_equal(n.numpy(), n2.numpy())
    cc = mo.get_cell_centers(p.numpy(), o.numpy(), n.numpy(), fp.numpy(), fo.numpy())
    assert np.array_equal(cc, hex_cell_centers(5, 4, 3))
