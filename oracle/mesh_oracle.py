"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of the two mesh-derived arrays the reference loader computes with Python loops
(/root/reference/openfoam_loader.py): `get_cell_centers` (:191-227) and `get_internal_cells` (:229-248).
It is *pinned*: tests/test_oracle_mesh.py checks it against `cell_centers` / `internal_mask` of the shipped
OpenFOAM case as the unmodified reference loader produced them here (tests/golden/shipped_mesh.npz and
shipped_polymesh.npz, written by oracle/make_golden.py).

One documented difference: the reference averages a cell's unique vertices in CPython `set` iteration order
(:216-220, `np.mean(points[list(set)], axis=0)` = sequential fp64 row adds, then a division); this restatement adds
them in ascending vertex id.  fp64 addition is not associative, so the two can differ in the last bits (measured on
the shipped case: 1.1e-16 absolute, 11040 of 147543 coordinates); the pin and the device parity gate are 1e-13 absolute, not bit equality.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np


def flatten_faces(faces):
    """faces as the loader returns them (`read_faces`: an object array / list of per-face vertex lists, ragged in
    general) -> (face_pts int32 [S], face_off int64 [F+1])."""
    lens = np.fromiter((len(f) for f in faces), dtype=np.int64, count=len(faces))
    off = np.zeros(len(faces) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    pts = np.empty(int(off[-1]), dtype=np.int32)
    for i, f in enumerate(faces):
        pts[off[i]:off[i + 1]] = np.asarray(f, dtype=np.int32)
    return pts, off


def num_cells(owner, neighbour) -> int:
    """openfoam_loader.py:197 / :236."""
    return int(max(np.max(owner), np.max(neighbour))) + 1


def get_cell_centers(points, owner, neighbour, face_pts, face_off) -> np.ndarray:
    """openfoam_loader.py:191-227.  Cell centre = mean of the unique vertices of the faces the cell owns (:203-207, face
    i of `owner[i]`) or neighbours (:210-214, face i of `neighbour[i]`); cells without a face stay (0,0,0) (:221-223).
    float64 [n_cells, 3]."""
    points = np.asarray(points, dtype=np.float64)
    owner = np.asarray(owner).astype(np.int64)
    neighbour = np.asarray(neighbour).astype(np.int64)
    face_off = np.asarray(face_off, dtype=np.int64)
    face_pts = np.asarray(face_pts).astype(np.int64)
    n_cells = num_cells(owner, neighbour)
    lens = np.diff(face_off)
    cells, pts = [], []
    for side in (owner, neighbour):
        nf = len(side)
        cells.append(np.repeat(side, lens[:nf]))
        pts.append(face_pts[:face_off[nf]])
    cells, pts = np.concatenate(cells), np.concatenate(pts)
    key = np.unique(cells * (int(points.shape[0]) + 1) + pts)          # sorted: by cell, then ascending vertex id
    kc, kp = key // (points.shape[0] + 1), key % (points.shape[0] + 1)
    out = np.zeros((n_cells, 3), dtype=np.float64)
    cnt = np.bincount(kc, minlength=n_cells)
    start = np.concatenate([[0], np.cumsum(cnt)])
    # sequential adds in ascending vertex order (np.add.at applies the updates in index order)
    np.add.at(out, kc, points[kp])
    nz = cnt > 0
    out[nz] /= cnt[nz, None]
    del start
    return out


def get_internal_cells(owner, neighbour) -> np.ndarray:
    """openfoam_loader.py:229-248: cells named by `neighbour`, and the owners of the first len(neighbour) faces."""
    owner = np.asarray(owner)
    neighbour = np.asarray(neighbour)
    mask = np.zeros(num_cells(owner, neighbour), dtype=bool)
    mask[neighbour] = True
    mask[owner[:len(neighbour)]] = True
    return mask
