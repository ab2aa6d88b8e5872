"""Device CSR cache.  PyG has no counterpart: GCNConv re-derives gcn_norm and GATConv re-does
remove/add self loops on EVERY forward of EVERY layer (cached=False at /root/reference/gnn_model.py:63).
Here the CSR of an `edge_index` is built once per distinct tensor (identity + version + shape) and
shared by all L layers of a forward (gnn_model.py:162-172 passes the same tensor to each layer) and
by the backward pass.  `batch.to(device)` (train.py:167) creates a fresh tensor each step, so in the
reference's loop the CSR is rebuilt once per step, not once per layer."""
from __future__ import annotations

import os
import weakref
from collections import OrderedDict

import torch

from . import ops


def identity_cached(store: dict, name: str, tensor: torch.Tensor, make):
    """make() cached in store[name] for as long as the SAME tensor object is passed with an unchanged version counter.
    Keyed on a weak reference to the object, not on its data pointer: the allocator hands a freed block to the next tensor
    of the same size, which would then be mistaken for the old one."""
    hit = store.get(name)
    if hit is not None and hit[0]() is tensor and hit[1] == tensor._version:
        return hit[2]
    value = make()
    store[name] = (weakref.ref(tensor), tensor._version, value)
    return value


class CSR:
    """One orientation of one variant: rowptr/col/eid (+dinv)."""
    __slots__ = ("rowptr", "col", "eid", "dinv", "nnz")

    def __init__(self, rowptr, col, eid, dinv):
        self.rowptr, self.col, self.eid, self.dinv = rowptr, col, eid, dinv
        self.nnz = col.numel()

    def pair(self):
        return (self.rowptr, self.col)


class Graph:
    """All CSRs derived from one edge_index [2,E] over N nodes (built lazily, cached).
    variant 'sl'  : self loops removed then one per node appended (GCNConv / GATConv)
    variant 'raw' : the list as given (GINConv / TransformerConv)."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError(f"edge_index must have shape [2, num_edges], got {tuple(edge_index.shape)}")
        ei = edge_index if edge_index.dtype == torch.int64 else edge_index.long()
        ei = ei.contiguous()
        # hold the caller's tensor weakly (a cached Graph must not pin a 1 GB edge_index of a finished
        # step); keep a strong reference only to a private contiguous/int64 copy
        self._ei_ref = weakref.ref(edge_index)
        self._ei_own = None if ei is edge_index else ei
        self.N = int(num_nodes)
        self.E = int(edge_index.shape[1])
        self._csr = {}
        self._perm = {}

    @property
    def edge_index(self) -> torch.Tensor:
        ei = self._ei_own if self._ei_own is not None else self._ei_ref()
        if ei is None:
            raise RuntimeError("b2g: the edge_index tensor of this cached graph was freed")
        return ei

    def csr(self, variant: str, by_source: bool = False) -> CSR:
        key = (variant, by_source)
        c = self._csr.get(key)
        if c is None:
            sl = variant == "sl"
            c = CSR(*ops.csr_build(self.edge_index, self.N, sl, by_source, want_dinv=(sl and not by_source)))
            self._csr[key] = c
        return c

    def max_degree(self, variant: str) -> int:
        """Largest row length of the target-major CSR (one device reduction per graph, cached)."""
        key = "_maxdeg_" + variant
        if not hasattr(self, key):
            rp = self.csr(variant, False).rowptr
            setattr(self, key, int((rp[1:] - rp[:-1]).max()) if self.N > 0 else 0)
        return getattr(self, key)

    def min_degree(self, variant: str) -> int:
        """Shortest row of the target-major CSR (one device reduction per graph, cached)."""
        key = "_mindeg_" + variant
        if not hasattr(self, key):
            rp = self.csr(variant, False).rowptr
            setattr(self, key, int((rp[1:] - rp[:-1]).min()) if self.N > 0 else 0)
        return getattr(self, key)

    def band(self) -> int:
        """max |source - target| over the edge list (one device reduction per graph, cached): the hint that lets
        the aggregation kernels sweep band-structured meshes panel by panel (aggregate.cu RowOrder)."""
        # Measured on B200 (cfg4, bf16 F=256): DRAM reads 9.6 -> 5.9-7.3 GB per launch, 3.00 -> 2.65 ms.
        # B2G_PANEL_ORDER=0 switches the hint off (linear sweep) for A/B runs.
        if os.environ.get("B2G_PANEL_ORDER", "1") == "0":
            return 0
        return self.band_owned()

    def band_raw(self) -> int:
        """max |source - target| (cached): rows [r0, r1) only read rows [r0 - band, r1 + band)."""
        if not hasattr(self, "_band"):
            ei = self.edge_index
            self._band = int((ei[0] - ei[1]).abs().max()) if ei.shape[1] else 0
        return self._band

    def band_owned(self) -> int:
        """The band over edges whose SOURCE is also a target-range node.  A partition's local graph numbers its ghost
        nodes after the owned ones (`distributed.py`): their |source - target| is ~N and would hide the band structure
        of the owned block from the row scheduler, although only a thin shell of rows reads ghosts."""
        if not hasattr(self, "_band_owned"):
            ei = self.edge_index
            if ei.shape[1] == 0:
                self._band_owned = 0
            else:
                n_tgt = ei[1].max() + 1
                d = (ei[0] - ei[1]).abs()
                self._band_owned = int(torch.where(ei[0] < n_tgt, d, torch.zeros_like(d)).max())
        return self._band_owned

    def dinv(self) -> torch.Tensor:
        """GCN deg^-1/2 over the self-loop-replaced list (in-degree by target)."""
        return self.csr("sl", False).dinv

    def edge_rows(self, variant: str, edge_attr: torch.Tensor) -> torch.Tensor:
        """Per-edge rows (e.g. edge_attr [E, 4]) as fp32 in the order of the target-major CSR of `variant`; cached for the
        tensor last passed (the reference hands the same edge_attr to every layer, gnn_model.py:170)."""
        if not edge_attr.is_cuda:
            raise RuntimeError("b2g: CUDA tensor required (this is the B200 path; there is no CPU fallback)")
        if variant == "sl":
            # GATConv(edge_dim): dropped loops lose their attributes, the new loops get the mean attribute of the row's other
            # entries (PyG fill_value='mean') — csrc/gat_fused.cu edge_rows_sl_kernel
            c = self.csr("sl", False)
            return identity_cached(self.__dict__, "_edge_rows_sl", edge_attr,
                                   lambda: ops.edge_rows_sl(edge_attr, c.eid, c.rowptr, self.N))
        return identity_cached(self.__dict__, "_edge_rows", edge_attr,
                               lambda: edge_attr.detach().float().index_select(0, self.csr("raw", False).eid.long()).contiguous())

    def perm(self, variant: str) -> torch.Tensor:
        """Position in the target-major CSR of each entry of the source-major CSR."""
        p = self._perm.get(variant)
        if p is None:
            a, b = self.csr(variant, False), self.csr(variant, True)
            p = ops.csr_perm(a.eid, b.eid, self.E + self.N)
            self._perm[variant] = p
        return p


_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_MAX = 8


def graph_of(edge_index: torch.Tensor, num_nodes: int) -> Graph:
    """Cached Graph for this edge_index tensor (keyed on storage identity, version, shape, N)."""
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), tuple(edge_index.stride()),
           int(num_nodes), edge_index.device.index)
    for k in [k for k, (r, _) in _CACHE.items() if r() is None]:   # drop graphs of freed tensors
        del _CACHE[k]
    hit = _CACHE.get(key)
    if hit is not None:
        ref, g = hit
        if ref() is edge_index:          # same live tensor object -> same content (version matched)
            _CACHE.move_to_end(key)
            return g
        del _CACHE[key]
    g = Graph(edge_index, num_nodes)
    _CACHE[key] = (weakref.ref(edge_index), g)
    if len(_CACHE) > _CACHE_MAX:
        # least recently used first; a pinned Graph (captured by a CUDA graph, or carrying a partition's patched ghost
        # deg^-1/2) stays for as long as its edge_index lives: a fresh copy would not be equivalent
        for k in [k for k, (_, gg) in _CACHE.items() if not getattr(gg, "_pinned", False)]:
            if len(_CACHE) <= _CACHE_MAX or k == key:
                break
            del _CACHE[k]
    return g


def clear_cache():
    _CACHE.clear()
