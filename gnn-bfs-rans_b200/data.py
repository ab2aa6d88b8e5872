"""Drop-in `torch_geometric.data.{Data,Batch}` (graph_constructor.py:8,262-267; train.py:9,155,167):
an attribute container with `.to(device)` and PyG's `Batch.from_data_list` concatenation rules."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.__dict__['_store'] = {}
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute bag semantics (the reference mutates graph.x / edge_index / edge_attr / num_nodes / y)
    def __getattr__(self, key):
        store = self.__dict__.get('_store', {})
        if key in store:
            return store[key]
        if key in ('x', 'edge_index', 'edge_attr', 'y', 'pos', 'batch'):
            return None
        if key == 'num_nodes':
            return self._infer_num_nodes()
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")

    def __setattr__(self, key, value):
        if value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __delattr__(self, key):
        self._store.pop(key, None)

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    def keys(self):
        return list(self._store.keys())

    def _infer_num_nodes(self):
        s = self._store
        if s.get('x') is not None:
            return s['x'].shape[0]
        if s.get('pos') is not None:
            return s['pos'].shape[0]
        if s.get('edge_index') is not None and s['edge_index'].numel() > 0:
            return int(s['edge_index'].max()) + 1
        return 0

    @property
    def num_edges(self):
        ei = self._store.get('edge_index')
        return 0 if ei is None else ei.shape[1]

    @property
    def num_node_features(self):
        x = self._store.get('x')
        return 0 if x is None else (1 if x.dim() == 1 else x.shape[1])

    def apply(self, fn):
        for k, v in list(self._store.items()):
            if isinstance(v, torch.Tensor):
                self._store[k] = fn(v)
        return self

    def to(self, device, non_blocking: bool = False):
        return self.apply(lambda t: t.to(device, non_blocking=non_blocking))

    def cpu(self):
        return self.to('cpu')

    def cuda(self, device=None, non_blocking: bool = False):
        return self.to('cuda' if device is None else device, non_blocking)

    def pin_memory(self):
        return self.apply(lambda t: t.pin_memory())

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out.__dict__['_store'] = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self._store.items()}
        return out

    def __repr__(self):
        parts = [f"{k}={list(v.shape) if isinstance(v, torch.Tensor) else v}" for k, v in self._store.items()]
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    """Batch.from_data_list: node-level tensors are concatenated along dim 0, `edge_index` along
    dim 1 with a cumulative node offset, plus `batch` (graph id per node) and `ptr`."""

    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        out = cls()
        if len(data_list) == 0:
            return out
        keys = data_list[0].keys()
        offs, n_tot = [], 0
        for d in data_list:
            offs.append(n_tot)
            n_tot += int(d.num_nodes)
        for k in keys:
            vals = [d[k] for d in data_list]
            if k == 'num_nodes':
                continue
            if not isinstance(vals[0], torch.Tensor):
                out._store[k] = vals if len(vals) > 1 else vals[0]
            elif k == 'edge_index' or k.endswith('_index'):
                out._store[k] = vals[0] if len(vals) == 1 else torch.cat([v + o for v, o in zip(vals, offs)], dim=1)
            elif vals[0].dim() == 0:
                out._store[k] = torch.stack(vals)
            else:
                out._store[k] = vals[0] if len(vals) == 1 else torch.cat(vals, dim=0)
        dev = next((v.device for v in out._store.values() if isinstance(v, torch.Tensor)), 'cpu')
        out._store['batch'] = torch.cat([torch.full((int(d.num_nodes),), i, dtype=torch.long, device=dev)
                                         for i, d in enumerate(data_list)])
        out._store['ptr'] = torch.tensor(offs + [n_tot], dtype=torch.long, device=dev)
        out._store['num_nodes'] = n_tot
        out.__dict__['_num_graphs'] = len(data_list)
        return out

    @property
    def num_graphs(self):
        return self.__dict__.get('_num_graphs', 1)
