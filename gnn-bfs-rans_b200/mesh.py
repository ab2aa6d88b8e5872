"""Device-side mesh ingest: the arrays `OpenFOAMLoader.load_mesh` derives from the polyMesh connectivity
with Python loops over every face (/root/reference/openfoam_loader.py:191-248) — minutes at 10 M cells — computed by
libb2g.so kernels (csrc/mesh.cu).  Same names, arguments and results as the loader's methods:

    get_cell_centers(points, owner, neighbour, faces)   -> [n_cells, 3] float64        (:191-227)
    get_internal_cells(owner, neighbour)                -> [n_cells] bool              (:229-248)
    derive_mesh(points, owner, neighbour, faces)        -> the dict entries `load_mesh` adds (:258-268)

`faces` is what `read_faces` returns (an object array / list of per-face vertex lists, ragged in general) or an already
flat pair `(face_pts int32 [S], face_off int64 [F+1])` — the form to use for big meshes, device tensors accepted.
The two functions return device tensors; `derive_mesh` returns the loader's numpy arrays by default, so that
`GraphConstructor({**parsed, **derive_mesh(...)})` is the reference's `GraphConstructor(loader.load_mesh())`.
Parsing the ASCII files stays with the loader (host I/O, out of scope)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import ops


def _device(device) -> torch.device:
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("b2g.mesh: no CUDA device (this is the B200 path; there is no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _i32(a, dev) -> torch.Tensor:
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype not in (torch.int32, torch.int64, torch.int16, torch.uint8, torch.int8):
        raise TypeError(f"b2g.mesh: integer ids expected, got {t.dtype}")
    return t.to(device=dev, dtype=torch.int32).contiguous()


def flatten_faces(faces) -> Tuple[np.ndarray, np.ndarray]:
    """`read_faces` output -> (face_pts int32 [S], face_off int64 [F+1]) on the host (a data-format conversion)."""
    if isinstance(faces, np.ndarray) and faces.ndim == 2:                      # all faces of one length (the shipped case)
        f, k = faces.shape
        return np.ascontiguousarray(faces.astype(np.int32)).reshape(-1), np.arange(f + 1, dtype=np.int64) * k
    lens = np.fromiter((len(f) for f in faces), dtype=np.int64, count=len(faces))
    off = np.zeros(len(faces) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    pts = np.fromiter((p for f in faces for p in f), dtype=np.int32, count=int(off[-1]))
    return pts, off


def _num_cells(o: torch.Tensor, n: torch.Tensor) -> int:
    if o.numel() == 0 or n.numel() == 0:                                       # np.max(owner) / np.max(neighbour), :197 / :236
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return ops.mesh_num_cells(o, n)


def _faces(faces, dev):
    if isinstance(faces, tuple) and len(faces) == 2:
        pts, off = faces
    else:
        pts, off = flatten_faces(faces)
    pts = _i32(pts, dev)
    off = (off if torch.is_tensor(off) else torch.from_numpy(np.ascontiguousarray(off))).to(device=dev, dtype=torch.int64).contiguous()
    if off.numel() < 1:
        raise ValueError("b2g.mesh: face_off must hold F + 1 offsets")
    return pts, off


def get_cell_centers(points, owner, neighbour, faces, device=None) -> torch.Tensor:
    """OpenFOAMLoader.get_cell_centers (openfoam_loader.py:191-227): centroid of the unique vertices of every cell."""
    dev = _device(device)
    pts = (points if torch.is_tensor(points) else torch.from_numpy(np.ascontiguousarray(np.asarray(points, dtype=np.float64))))
    pts = pts.to(device=dev, dtype=torch.float64).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError("points must be [P, 3]")
    o, n = _i32(owner, dev), _i32(neighbour, dev)
    fp, fo = _faces(faces, dev)
    n_faces = fo.numel() - 1
    if o.numel() > n_faces or n.numel() > n_faces:
        raise IndexError("b2g.mesh: owner / neighbour name more faces than `faces` holds")        # faces[i], :205 / :212
    n_cells = _num_cells(o, n)
    ends = fo[torch.tensor([o.numel(), n.numel()], device=dev)].tolist()                         # one small host read
    return ops.mesh_cell_centers(pts, o, n, fp, fo, n_cells, int(ends[0]) + int(ends[1]))


def get_internal_cells(owner, neighbour, device=None) -> torch.Tensor:
    """OpenFOAMLoader.get_internal_cells (openfoam_loader.py:229-248)."""
    dev = _device(device)
    o, n = _i32(owner, dev), _i32(neighbour, dev)
    if n.numel() > o.numel():
        raise IndexError("b2g.mesh: neighbour longer than owner")                                  # owner[i], :244
    return ops.mesh_internal_cells(o, n, _num_cells(o, n))


def derive_mesh(points, owner, neighbour, faces, device=None, as_numpy: bool = True) -> Dict[str, object]:
    """The entries `load_mesh` computes rather than parses (openfoam_loader.py:255-268): cell_centers, n_cells,
    internal_mask, n_internal_cells.  as_numpy=False keeps the two arrays on the device."""
    cc = get_cell_centers(points, owner, neighbour, faces, device)
    mask = get_internal_cells(owner, neighbour, cc.device)
    n_int = int(mask.sum())
    if as_numpy:
        cc, mask = cc.cpu().numpy(), mask.cpu().numpy()
    return {"cell_centers": cc, "n_cells": int(cc.shape[0]), "internal_mask": mask, "n_internal_cells": n_int}
