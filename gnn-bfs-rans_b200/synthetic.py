"""Synthetic meshes of BASELINE.json's configs, emitted the way OpenFOAM stores them: owner/neighbour
face lists with owner < neighbour, internal faces sorted by owner (upper-triangular order), boundary
faces after them."""
from __future__ import annotations

import numpy as np
import torch


def hex_mesh_faces(nx: int, ny: int, nz: int, device="cpu", boundary: bool = False):
    """Structured hex block, lexicographic cell ids (x fastest).  Returns int32 (owner, neighbour):
    internal faces only unless boundary=True (then owner also lists one entry per boundary face)."""
    dev = torch.device(device)
    N = nx * ny * nz
    ids = torch.arange(N, dtype=torch.int64, device=dev)
    ix = ids % nx
    iy = (ids // nx) % ny
    iz = ids // (nx * ny)
    # per cell up to three "upper" faces (+x, +y, +z), interleaved per owner so the list is sorted by owner
    has = torch.stack([ix < nx - 1, iy < ny - 1, iz < nz - 1], dim=1)
    nb = torch.stack([ids + 1, ids + nx, ids + nx * ny], dim=1)
    own = ids.unsqueeze(1).expand(-1, 3)
    owner = own[has].to(torch.int32)
    nei = nb[has].to(torch.int32)
    if boundary:
        nbnd = (ix == 0).int() + (ix == nx - 1).int() + (iy == 0).int() + (iy == ny - 1).int() + \
               (iz == 0).int() + (iz == nz - 1).int()
        owner = torch.cat([owner, torch.repeat_interleave(ids, nbnd.long()).to(torch.int32)])
    return owner.contiguous(), nei.contiguous()


def hex_polymesh(nx: int, ny: int, nz: int, device="cpu"):
    """The same block as a polyMesh: (points fp64 [P,3], owner int32 [F], neighbour int32 [F_int], face_pts int32 [4F],
    face_off int64 [F+1]).  Internal faces first, in hex_mesh_faces order (+x, +y, +z per owner), then the boundary faces
    in owner order (-x, +x, -y, +y, -z, +z).  Vertex id = px + (nx+1) (py + (ny+1) pz), coordinates = (px, py, pz), so a
    cell's centre is exactly (ix + .5, iy + .5, iz + .5) in fp64."""
    dev = torch.device(device)
    N = nx * ny * nz
    ids = torch.arange(N, dtype=torch.int64, device=dev)
    ix, iy, iz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    v0 = ix * sx + iy * sy + iz * sz                                   # the cell's (0,0,0) corner

    def quad(axis_off, a, b):                                           # 4 corners of the face at offset axis_off
        return torch.stack([v0 + axis_off, v0 + axis_off + a, v0 + axis_off + a + b, v0 + axis_off + b], dim=1)

    qx0, qx1 = quad(0, sy, sz), quad(sx, sy, sz)
    qy0, qy1 = quad(0, sx, sz), quad(sy, sx, sz)
    qz0, qz1 = quad(0, sx, sy), quad(sz, sx, sy)
    has = torch.stack([ix < nx - 1, iy < ny - 1, iz < nz - 1], dim=1)
    nb = torch.stack([ids + 1, ids + nx, ids + nx * ny], dim=1)
    own = ids.unsqueeze(1).expand(-1, 3)
    q_int = torch.stack([qx1, qy1, qz1], dim=1)[has]                   # [F_int, 4]
    bnd = torch.stack([ix == 0, ix == nx - 1, iy == 0, iy == ny - 1, iz == 0, iz == nz - 1], dim=1)
    q_bnd = torch.stack([qx0, qx1, qy0, qy1, qz0, qz1], dim=1)[bnd]
    owner = torch.cat([own[has], ids.unsqueeze(1).expand(-1, 6)[bnd]]).to(torch.int32).contiguous()
    neighbour = nb[has].to(torch.int32).contiguous()
    face_pts = torch.cat([q_int, q_bnd]).to(torch.int32).reshape(-1).contiguous()
    face_off = torch.arange(owner.numel() + 1, dtype=torch.int64, device=dev) * 4
    P = (nx + 1) * (ny + 1) * (nz + 1)
    pid = torch.arange(P, dtype=torch.int64, device=dev)
    points = torch.stack([pid % (nx + 1), (pid // (nx + 1)) % (ny + 1), pid // sz], dim=1).to(torch.float64).contiguous()
    return points, owner, neighbour, face_pts, face_off


def hex_cell_centers(nx: int, ny: int, nz: int) -> np.ndarray:
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([x.ravel() + 0.5, y.ravel() + 0.5, z.ravel() + 0.5], axis=1).astype(np.float64)


def delaunay_dual_faces(n_points: int, seed: int = 1):
    """cfg3: cells = triangles of a Delaunay triangulation of uniform points in the unit square;
    faces = shared triangle edges.  Returns (owner, neighbour int32 numpy, centres float64 [N,3])."""
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(seed).random((n_points, 2))
    tri = Delaunay(pts)
    nbr = tri.neighbors                       # [N,3], -1 at the hull
    N = len(nbr)
    own = np.repeat(np.arange(N), 3)
    nb = nbr.ravel()
    keep = (nb >= 0) & (own < nb)             # each shared edge once, owner < neighbour
    own, nb = own[keep], nb[keep]
    order = np.lexsort((nb, own))
    c = pts[tri.simplices].mean(axis=1)
    centres = np.concatenate([c, np.zeros((N, 1))], axis=1)
    return own[order].astype(np.int32), nb[order].astype(np.int32), centres


def hilbert_index_2d(x: np.ndarray, y: np.ndarray, bits: int = 16) -> np.ndarray:
    """Distance along the 2-D Hilbert curve of order `bits` for integer coordinates in [0, 2^bits) (the classic
    xy -> d walk, vectorised over the points)."""
    n = np.uint64(1) << np.uint64(bits)
    x = x.astype(np.uint64).copy()
    y = y.astype(np.uint64).copy()
    d = np.zeros(x.shape, dtype=np.uint64)
    s = n >> np.uint64(1)
    while s > 0:
        rx = (x & s) > 0
        ry = (y & s) > 0
        d += s * s * ((np.uint64(3) * rx.astype(np.uint64)) ^ ry.astype(np.uint64))
        flip = (~ry) & rx                                 # rotate the quadrant so the sub-curve has the standard orientation
        x = np.where(flip, n - np.uint64(1) - x, x)
        y = np.where(flip, n - np.uint64(1) - y, y)
        x, y = np.where(~ry, y, x), np.where(~ry, x, y)
        s >>= np.uint64(1)
    return d


def hilbert_renumber_2d(owner: np.ndarray, neighbour: np.ndarray, centres: np.ndarray, bits: int = 16):
    """Renumber the cells of a 2-D mesh along the Hilbert curve through their centres (cfg3's "Hilbert-sorted variant",
    SURVEY §8d).  Returns (owner, neighbour, centres) of the renumbered mesh, again with owner < neighbour and the faces
    sorted by (owner, neighbour), plus `order` (new id -> old id)."""
    c = centres[:, :2]
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = np.minimum(((c - lo) / np.maximum(hi - lo, 1e-300) * (2 ** bits)).astype(np.int64), 2 ** bits - 1)
    order = np.argsort(hilbert_index_2d(q[:, 0], q[:, 1], bits), kind="stable")
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    a, b = rank[owner], rank[neighbour]
    o2, n2 = np.minimum(a, b), np.maximum(a, b)
    f = np.lexsort((n2, o2))
    return o2[f].astype(np.int32), n2[f].astype(np.int32), centres[order], order
