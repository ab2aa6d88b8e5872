"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Vectorised numpy restatement of the reference graph builder
(/root/reference/graph_constructor.py).  It is *pinned*: tests/test_oracle_builder.py checks it
against the toy known-answer vectors and the SHA-256 goldens of the shipped OpenFOAM case that
oracle/make_golden.py produced by running the unmodified reference in this container.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np


def build_edge_index(owner: np.ndarray, neighbour: np.ndarray) -> np.ndarray:
    """graph_constructor.py:28-56.  Internal face i -> edges (o_i,n_i),(n_i,o_i) interleaved;
    every face past len(neighbour) -> one (o,o) self-loop (duplicates kept).  int64 [2,E]."""
    owner = np.asarray(owner)
    neighbour = np.asarray(neighbour)
    n_int = len(neighbour)
    if n_int > len(owner):
        raise IndexError("neighbour longer than owner")  # reference: owner[i] IndexError, :40
    o = owner[:n_int].astype(np.int64)
    n = neighbour.astype(np.int64)
    src = np.empty(2 * n_int, dtype=np.int64)
    dst = np.empty(2 * n_int, dtype=np.int64)
    src[0::2], dst[0::2] = o, n          # :44
    src[1::2], dst[1::2] = n, o          # :45
    b = owner[n_int:].astype(np.int64)   # :49-53
    return np.stack([np.concatenate([src, b]), np.concatenate([dst, b])])


def _edge_attr(cell_centers: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """graph_constructor.py:190-219 (inline copy of :58-90).  float64 maths, cast to fp32 at the end
    (torch.tensor(list_of_float64, dtype=float32), :219)."""
    E = edge_index.shape[1]
    out = np.zeros((E, 4), dtype=np.float64)
    if E == 0:
        return out.astype(np.float32)
    cc = np.asarray(cell_centers, dtype=np.float64)
    src, dst = edge_index[0], edge_index[1]
    n = len(cc)
    ok = (src >= 0) & (src < n) & (dst >= 0) & (dst < n) & (src != dst)   # :203, :208
    d = cc[dst[ok]] - cc[src[ok]]                                           # :213
    # np.linalg.norm of a 3-vector = sqrt(dot(d,d)); summation order x,y,z        :214
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    pos = dist > 0
    dn = d.copy()
    dn[pos] = d[pos] / dist[pos, None]                                      # :215-216
    out[ok, :3] = dn
    out[ok, 3] = dist
    return out.astype(np.float32)


def compute_edge_attributes(cell_centers: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """graph_constructor.py:58-90 (no range check there: out-of-range ids raise IndexError)."""
    cc = np.asarray(cell_centers)
    ne = edge_index[:, edge_index[0] != edge_index[1]]
    if ne.size and (ne.max() >= len(cc) or ne.min() < 0):
        raise IndexError("edge endpoint outside cell_centers")   # :79-80 would raise (or wrap)
    return _edge_attr(cc, edge_index)


def build_graph(mesh_data: dict, field_data=None, node_features=None,
                filter_internal: bool = False, n_internal_cells=None) -> dict:
    """graph_constructor.py:92-269.  Returns dict(x fp32[N,*], edge_index int64[2,E],
    edge_attr fp32[E,4], num_nodes int)."""
    owner = np.asarray(mesh_data['owner'])
    neighbour = np.asarray(mesh_data['neighbour'])
    cell_centers = np.asarray(mesh_data['cell_centers'])
    n_cells = int(mesh_data['n_cells'])

    internal_mask = None
    if filter_internal:                                                     # :109-129
        if n_internal_cells is not None:
            n_nodes = int(n_internal_cells)
            internal_indices = np.arange(n_nodes)
            internal_mask = np.zeros(n_cells, dtype=bool)
            internal_mask[:n_nodes] = True
        elif 'internal_mask' in mesh_data:
            internal_mask = np.asarray(mesh_data['internal_mask']).astype(bool)
            internal_indices = np.where(internal_mask)[0]
            n_nodes = len(internal_indices)
        else:                                                               # :120-125
            internal_indices = np.arange(n_cells)
            n_nodes = n_cells
            filter_internal = False
        if filter_internal:
            old_to_new = np.full(n_cells, -1, dtype=np.int32)
            old_to_new[internal_indices] = np.arange(n_nodes)              # IndexError if n > n_cells
    if not filter_internal:                                                 # :130-134
        internal_indices = np.arange(n_cells)
        n_nodes = n_cells

    if filter_internal and internal_mask is not None:                       # :137-154
        n_int = len(neighbour)
        o = owner[:n_int].astype(np.int64)
        n = neighbour.astype(np.int64)
        keep = internal_mask[o] & internal_mask[n]                          # IndexError if id >= n_cells
        no = old_to_new[o[keep]].astype(np.int64)
        nn = old_to_new[n[keep]].astype(np.int64)
        src = np.empty(2 * len(no), dtype=np.int64)
        dst = np.empty(2 * len(no), dtype=np.int64)
        src[0::2], dst[0::2] = no, nn
        src[1::2], dst[1::2] = nn, no
        edge_index = np.stack([src, dst])
    else:
        edge_index = build_edge_index(owner, neighbour)                     # :156

    if edge_index.shape[1] > 0:                                             # :168-173
        if edge_index.max() >= n_nodes:
            valid = (edge_index[0] < n_nodes) & (edge_index[1] < n_nodes)
            edge_index = edge_index[:, valid]

    if edge_index.shape[1] > 0 and n_nodes > 0:                             # :176-187
        connected = np.unique(edge_index.ravel())
        all_nodes = np.arange(n_nodes)
        isolated = all_nodes[~np.isin(all_nodes, connected)]
        if len(isolated) > 0:
            edge_index = np.concatenate([edge_index, np.stack([isolated, isolated])], axis=1)

    if edge_index.shape[1] > 0:                                             # :190-219
        cc = cell_centers[internal_indices] if (filter_internal and internal_mask is not None) else cell_centers
        edge_attr = _edge_attr(cc, edge_index)
    else:                                                                   # :220-227
        if n_nodes > 0:
            edge_index = np.tile(np.arange(n_nodes, dtype=np.int64), (2, 1))
            edge_attr = np.zeros((n_nodes, 4), dtype=np.float32)
        else:
            edge_attr = np.zeros((0, 4), dtype=np.float32)

    if node_features is None:                                               # :230-239
        nf = cell_centers
    else:
        nf = np.asarray(node_features)
    if filter_internal and internal_mask is not None:
        nf = nf[internal_indices].copy()
    else:
        nf = nf.copy()
    if field_data is not None:                                              # :242-256
        feats = [nf]
        if 'U' in field_data:
            feats.append(field_data['U'])
        for name in ['p', 'k', 'epsilon', 'nut']:
            if name in field_data:
                feats.append(np.asarray(field_data[name]).reshape(-1, 1))
        nf = np.hstack(feats)
    x = nf.astype(np.float32)                                               # :259
    return dict(x=x, edge_index=edge_index.astype(np.int64), edge_attr=edge_attr, num_nodes=n_nodes)


def get_boundary_mask(mesh_data: dict, boundary_name: str) -> np.ndarray:
    """graph_constructor.py:271-295."""
    if boundary_name not in mesh_data['boundaries']:
        raise ValueError(f"Boundary {boundary_name} not found")
    info = mesh_data['boundaries'][boundary_name]
    owner = np.asarray(mesh_data['owner'])
    s, n = info['startFace'], info['nFaces']
    mask = np.zeros(int(mesh_data['n_cells']), dtype=bool)
    idx = owner[s:min(s + n, len(owner))]
    mask[idx] = True
    return mask


def effective_edges(edge_index: np.ndarray, num_nodes: int, self_loops_replaced: bool) -> np.ndarray:
    """The edge list a layer actually aggregates over: PyG remove_self_loops +
    add_self_loops appends arange(N) pairs at the END (GCN/GAT); GIN/Transformer use the raw list."""
    ei = np.asarray(edge_index)
    if not self_loops_replaced:
        return ei
    keep = ei[0] != ei[1]
    loops = np.arange(num_nodes, dtype=ei.dtype)
    return np.concatenate([ei[:, keep], np.stack([loops, loops])], axis=1)


def csr_by_target(edge_index: np.ndarray, num_nodes: int, by_source: bool = False):
    """CSR definition (no reference counterpart): STABLE sort of the edge list by target
    (or by source for the transposed CSR); col = the other endpoint; eid = position in the edge list."""
    ei = np.asarray(edge_index)
    key = ei[0] if by_source else ei[1]
    other = ei[1] if by_source else ei[0]
    order = np.argsort(key, kind='stable')
    counts = np.bincount(key, minlength=num_nodes)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr.astype(np.int32), other[order].astype(np.int32), order.astype(np.int32)
