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
    dim 1 with a cumulative node offset, plus `batch` (graph id per node) and `ptr`.

    Device batching.  The reference collates on the host and then moves the batch (train.py:155,167:
    `Batch.from_data_list(batch)` ... `batch.to(device)`).  Here the concatenation is DEFERRED: from_data_list keeps the
    sample list, and `.to(cuda)` copies every sample's tensors straight into their slices of the batch tensors on the device
    (no host-side torch.cat pass over the data) and finishes with one kernel that adds the node offsets to edge_index and
    writes the `batch` vector (csrc/train_glue.cu batch_finalize_kernel).  Any attribute access before `.to()` materialises
    the batch on the host exactly as before, so the semantics are unchanged."""

    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        out = cls()
        if len(data_list) == 0:
            return out
        if cls._deferrable(data_list):
            out.__dict__['_pending'] = list(data_list)
            out.__dict__['_num_graphs'] = len(data_list)
            return out
        return cls._collate_host(out, data_list)

    # ---------------------------------------------------------------- deferred (device) collation
    @staticmethod
    def _deferrable(data_list) -> bool:
        keys = data_list[0].keys()
        for d in data_list:
            if d.keys() != keys:
                return False
            for k in keys:
                v = d[k]
                if k == 'num_nodes':
                    continue
                if not isinstance(v, torch.Tensor) or v.is_cuda or v.dim() == 0:
                    return False
                if (k == 'edge_index' or k.endswith('_index')) and (v.dim() != 2 or v.shape[0] != 2 or v.dtype != torch.int64):
                    return False
        return 'edge_index' in keys

    def _materialize(self):
        pend = self.__dict__.pop('_pending', None)
        if pend is not None:
            Batch._collate_host(self, pend)

    def __getattr__(self, key):
        if key not in ('_store', '_pending', '_num_graphs') and '_pending' in self.__dict__:
            self._materialize()
        return super().__getattr__(key)

    def __setattr__(self, key, value):
        self._materialize()
        super().__setattr__(key, value)

    def __setitem__(self, key, value):
        self._materialize()
        super().__setitem__(key, value)

    def __getitem__(self, key):
        self._materialize()
        return super().__getitem__(key)

    def __contains__(self, key):
        self._materialize()
        return super().__contains__(key)

    def keys(self):
        self._materialize()
        return super().keys()

    def apply(self, fn):
        self._materialize()
        return super().apply(fn)

    def clone(self):
        self._materialize()
        return super().clone()

    def __repr__(self):
        self._materialize()
        return super().__repr__()

    def to(self, device, non_blocking: bool = False):
        pend = self.__dict__.get('_pending')
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if pend is None or dev.type != 'cuda':
            return super().to(device, non_blocking=non_blocking)
        self.__dict__.pop('_pending')
        self._collate_device(pend, dev)
        return self

    def _collate_device(self, data_list, dev):
        from . import _lib
        lib = _lib.load()
        keys = data_list[0].keys()
        node_ptr, edge_ptr = [0], [0]
        for d in data_list:
            node_ptr.append(node_ptr[-1] + int(d.num_nodes))
            edge_ptr.append(edge_ptr[-1] + int(d['edge_index'].shape[1]))
        n_tot, e_tot = node_ptr[-1], edge_ptr[-1]
        for k in keys:
            if k == 'num_nodes':
                continue
            vals = [d[k] for d in data_list]
            if k == 'edge_index' or k.endswith('_index'):
                out = torch.empty((2, sum(v.shape[1] for v in vals)), dtype=torch.int64, device=dev)
                off = 0
                for v in vals:
                    out[:, off:off + v.shape[1]].copy_(v, non_blocking=True)
                    off += v.shape[1]
            else:
                out = torch.empty((sum(v.shape[0] for v in vals),) + tuple(vals[0].shape[1:]), dtype=vals[0].dtype, device=dev)
                off = 0
                for v in vals:
                    out[off:off + v.shape[0]].copy_(v, non_blocking=True)
                    off += v.shape[0]
            self._store[k] = out
        ptrs = torch.tensor([edge_ptr, node_ptr], dtype=torch.int64).to(dev, non_blocking=True)
        batch = torch.empty(n_tot, dtype=torch.int64, device=dev)
        ei = self._store['edge_index']
        _lib.check(lib.b2g_batch_finalize(ei.data_ptr(), e_tot, ptrs[0].data_ptr(), ptrs[1].data_ptr(), len(data_list),
                                          batch.data_ptr(), n_tot, torch.cuda.current_stream(dev).cuda_stream), "batch_finalize")
        for k in keys:                                     # other *_index attributes get the same node offsets
            if k != 'edge_index' and k.endswith('_index'):
                t = self._store[k]
                _lib.check(lib.b2g_batch_finalize(t.data_ptr(), t.shape[1], ptrs[0].data_ptr(), ptrs[1].data_ptr(),
                                                  len(data_list), None, 0, torch.cuda.current_stream(dev).cuda_stream),
                           "batch_finalize")
        self._store['batch'] = batch
        self._store['ptr'] = ptrs[1]
        self._store['num_nodes'] = n_tot
        self.__dict__['_num_graphs'] = len(data_list)

    # ---------------------------------------------------------------- host collation (PyG semantics)
    @staticmethod
    def _collate_host(out, data_list):
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
