"""Drop-in for /root/reference/graph_constructor.py: same class, methods and keyword arguments,
bit-exact outputs, but the O(faces) / O(edges) Python loops (:39-55, :140-154, :178-187, :198-217)
run as libb2g.so kernels (K0 builder, K0c edge attributes) on the GPU.

Like the reference it returns CPU tensors inside a `Data` (callers index graph.edge_index on the
host, train.py:114, and move batches with .to(device), train.py:167).  `build_graph(...,
device='cuda')` is the extension that keeps everything resident for meshes where a host round trip
of a 1 GB edge_index would dominate."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .data import Data


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b2g GraphConstructor: a CUDA device is required (no CPU fallback)")
    return torch.device('cuda' if device is None or str(device) == 'cpu' else device)


class GraphConstructor:
    """Constructs graph from OpenFOAM mesh data (graph_constructor.py:12-26)."""

    def __init__(self, mesh_data: Dict):
        self.mesh_data = mesh_data
        self.owner = mesh_data['owner']
        self.neighbour = mesh_data['neighbour']
        self.cell_centers = mesh_data['cell_centers']
        self.n_cells = mesh_data['n_cells']
        self._dev_faces = None

    # ------------------------------------------------------------------ helpers
    def _faces(self, device=None):
        dev = _dev(device)
        if self._dev_faces is None or self._dev_faces[0].device != dev:
            o = torch.as_tensor(np.ascontiguousarray(self.owner, dtype=np.int32)).to(dev)
            n = torch.as_tensor(np.ascontiguousarray(self.neighbour, dtype=np.int32)).to(dev)
            self._dev_faces = (o, n)
        return self._dev_faces

    # ------------------------------------------------------------------ graph_constructor.py:28-56
    def build_edge_index(self, device=None) -> torch.Tensor:
        o, n = self._faces(device)
        ei = ops.build_edge_index(o, n)
        return ei if device is not None and str(device) != 'cpu' else ei.cpu()

    # ------------------------------------------------------------------ graph_constructor.py:58-90
    def compute_edge_attributes(self, edge_index: torch.Tensor, device=None) -> torch.Tensor:
        cc = np.asarray(self.cell_centers, dtype=np.float64).reshape(-1, 3)
        if edge_index.shape[1] > 0:
            nl = edge_index[:, edge_index[0] != edge_index[1]]
            if nl.numel() and (int(nl.max()) >= len(cc) or int(nl.min()) < 0):
                raise IndexError("index out of bounds for cell_centers")       # reference: :79-80
        dev = _dev(device if device is not None else (edge_index.device if edge_index.is_cuda else None))
        ea = ops.edge_attr(torch.from_numpy(np.ascontiguousarray(cc)).to(dev), edge_index.to(dev))
        return ea if edge_index.is_cuda or (device is not None and str(device) != 'cpu') else ea.cpu()

    # ------------------------------------------------------------------ graph_constructor.py:92-269
    def build_graph(self, field_data: Optional[Dict] = None, node_features: Optional[np.ndarray] = None,
                    filter_internal: bool = False, n_internal_cells: Optional[int] = None, device=None) -> Data:
        keep_on_device = device is not None and str(device) != 'cpu'
        dev = _dev(device)
        n_cells = int(self.n_cells)
        o, n = self._faces(dev)
        internal_indices = None
        mode, o2n = 0, None
        if filter_internal:
            if n_internal_cells is not None:                                   # :110-115
                n_nodes = int(n_internal_cells)
                if n_nodes > n_cells:
                    raise IndexError(f"index {n_cells} is out of bounds for axis 0 with size {n_cells}")  # :129
                internal_indices = slice(0, n_nodes)
                mode = 1
            elif 'internal_mask' in self.mesh_data:                            # :116-119
                mask = np.asarray(self.mesh_data['internal_mask']).astype(bool)
                if len(mask) != n_cells:
                    raise IndexError("internal_mask length does not match n_cells")
                o2n, n_nodes = ops.mask_to_map(torch.from_numpy(mask.view(np.uint8)).to(dev))
                internal_indices = np.where(mask)[0]
                mode = 1
            else:                                                              # :120-125
                n_nodes = n_cells
        else:
            n_nodes = n_cells                                                  # :130-134

        edge_index = ops.build_graph_edges(o, n, mode, o2n, n_cells, n_nodes)  # :137-187, 220-227

        cc = np.asarray(self.cell_centers, dtype=np.float64).reshape(-1, 3)
        cc_used = cc[internal_indices] if internal_indices is not None else cc  # :192-195
        if edge_index.shape[1] > 0:
            edge_attr = ops.edge_attr(torch.from_numpy(np.ascontiguousarray(cc_used)).to(dev), edge_index)
        else:
            edge_attr = torch.empty((0, 4), dtype=torch.float32, device=dev)  # :227

        if node_features is None:                                              # :230-239
            nf = np.asarray(self.cell_centers)
        else:
            nf = np.asarray(node_features)
        nf = nf[internal_indices].copy() if internal_indices is not None else nf.copy()
        if field_data is not None:                                             # :242-256
            feats = [nf]
            if 'U' in field_data:
                feats.append(field_data['U'])
            for name in ['p', 'k', 'epsilon', 'nut']:
                if name in field_data:
                    feats.append(np.asarray(field_data[name]).reshape(-1, 1))
            nf = np.hstack(feats)
        x = torch.tensor(nf, dtype=torch.float32)                              # :259
        if keep_on_device:
            return Data(x=x.to(dev), edge_index=edge_index, edge_attr=edge_attr, num_nodes=n_nodes)
        return Data(x=x, edge_index=edge_index.cpu(), edge_attr=edge_attr.cpu(), num_nodes=n_nodes)

    # ------------------------------------------------------------------ graph_constructor.py:271-295
    def get_boundary_mask(self, boundary_name: str) -> np.ndarray:
        if boundary_name not in self.mesh_data['boundaries']:
            raise ValueError(f"Boundary {boundary_name} not found")
        info = self.mesh_data['boundaries'][boundary_name]
        s, nf = info['startFace'], info['nFaces']
        mask = np.zeros(self.n_cells, dtype=bool)
        owner = np.asarray(self.owner)
        mask[owner[s:min(s + nf, len(owner))]] = True
        return mask
