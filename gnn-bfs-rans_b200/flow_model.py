"""Host-side mirror of the reference's model wrapper `FlowGNN` (/root/reference/gnn_model.py:14-220),
the CALLER of the hot path.  The reference's own file runs unchanged through dropin.install(); this
mirror exists because /root/reference is not present on the GPU box, and bench.py / the GPU tests
need the same caller there: same constructor arguments, same submodule names (so state_dicts are
interchangeable: input_proj, gnn_layers.N, batch_norms.N, output_proj.{0,3,6,8}), same forward
order (layer -> residual -> BatchNorm -> ReLU -> dropout, then the 4-Linear head)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as tnn

from . import functional as Fn
from . import nn as gnn
from . import ops

LAYER_TYPES = ('GCN', 'GAT', 'GIN', 'Transformer')


def _make_layer(layer_type: str, width: int, dropout: float, edge_dim=None):
    if layer_type == 'GCN':
        return gnn.GCNConv(width, width)                                           # gnn_model.py:63
    if layer_type == 'GAT':
        return gnn.GATConv(width, width, heads=4, concat=False, dropout=dropout, edge_dim=edge_dim)   # :65-68
    if layer_type == 'GIN':
        return gnn.GINConv(tnn.Sequential(tnn.Linear(width, width), tnn.ReLU(), tnn.Linear(width, width)))  # :70-75
    if layer_type == 'Transformer':
        return gnn.TransformerConv(width, width, heads=4, concat=False, dropout=dropout, edge_dim=edge_dim)  # :77-80
    raise ValueError(f"Unknown layer type: {layer_type}")                         # :82


class FlowGNN(tnn.Module):
    def __init__(self, input_dim: int = 3, hidden_dim: int = 128, output_dim: int = 8, num_layers: int = 4,
                 layer_type: str = 'GCN', use_edge_attr: bool = True, dropout: float = 0.1,
                 use_batch_norm: bool = True, validate_edges: bool = True, fused_glue: bool = False,
                 edge_dim: Optional[int] = None):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_layers, self.layer_type = num_layers, layer_type
        self.use_edge_attr, self.use_batch_norm = use_edge_attr, use_batch_norm
        self.validate_edges = validate_edges     # the reference's two .item() syncs per forward (:131-132)
        # opt-in: residual add + BatchNorm + ReLU + dropout (:184-192) as the two passes of csrc/bn.cu
        # instead of four torch ops.  Same arithmetic; the dropout mask comes from the library's Philox stream.
        self.fused_glue = fused_glue
        # opt-in: edge_dim=4 builds the Transformer (or GAT) layers with PyG's lin_edge so that the edge_attr the
        # reference already passes (:170) is consumed — attention over edge features (THEORY_AND_METHODS.md:165-166).
        # None = the reference's constructor call (:77-80), where the attribute cannot be used.
        if edge_dim is not None and layer_type not in ('Transformer', 'GAT'):
            raise ValueError("edge_dim is an option of layer_type='Transformer' / 'GAT'")
        self.edge_dim = edge_dim
        self.input_proj = tnn.Linear(input_dim, hidden_dim)
        self.gnn_layers = tnn.ModuleList(_make_layer(layer_type, hidden_dim, dropout, edge_dim) for _ in range(num_layers))
        self.batch_norms = tnn.ModuleList(gnn.BatchNorm(hidden_dim) for _ in range(num_layers)) if use_batch_norm else None
        h = hidden_dim
        self.output_proj = tnn.Sequential(
            tnn.Linear(h, h), tnn.ReLU(), tnn.Dropout(dropout),
            tnn.Linear(h, h), tnn.ReLU(), tnn.Dropout(dropout),
            tnn.Linear(h, h // 2), tnn.ReLU(),
            tnn.Linear(h // 2, output_dim))
        self.dropout = tnn.Dropout(dropout)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None,
                batch: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = x.shape[0]
        if edge_index.shape[0] != 2:
            raise ValueError(f"edge_index must have shape [2, num_edges], got {edge_index.shape}")
        if self.validate_edges and edge_index.shape[1] > 0:                        # :130-149
            lo, hi = int(edge_index.min()), int(edge_index.max())
            if lo < 0 or hi >= n:
                ok = ((edge_index >= 0) & (edge_index < n)).all(dim=0)
                edge_index = edge_index[:, ok]
                if edge_attr is not None and edge_attr.shape[0] > 0:
                    edge_attr = edge_attr[ok]
            if edge_index.shape[1] == 0:
                edge_index = torch.arange(n, dtype=torch.long, device=x.device).repeat(2, 1)
                if edge_attr is not None:
                    edge_attr = edge_attr.new_zeros((n, edge_attr.shape[1]))
        if edge_attr is not None and edge_index.shape[1] > 0 and edge_attr.shape[0] != edge_index.shape[1]:
            raise ValueError(f"edge_attr must have {edge_index.shape[1]} entries, got {edge_attr.shape[0]}")
        h = self.input_proj(x)
        for i, layer in enumerate(self.gnn_layers):
            try:
                if self.layer_type == 'Transformer' or (self.layer_type == 'GAT' and self.edge_dim is not None):
                    h_new = layer(h, edge_index, edge_attr=edge_attr)              # :170 (GAT: only with the edge_dim opt-in)
                else:
                    h_new = layer(h, edge_index)                                   # :166,168
            except RuntimeError as e:                                              # :173-181
                raise RuntimeError(f"Message passing failed in layer {i} ({self.layer_type}): {e}\n"
                                   f"  num_nodes: {n}, num_edges: {edge_index.shape[1]}, x shape: {tuple(h.shape)}") from e
            if self.fused_glue and self.use_batch_norm and ops.bn_supported(h):
                h = Fn.batch_norm(h, h_new, self.batch_norms[i].module, relu=True, p_drop=self.dropout.p)
                continue
            h = h + h_new
            if self.use_batch_norm:
                h = self.batch_norms[i](h)
            h = self.dropout(torch.relu(h))
        return self.output_proj(h)

    def predict_fields(self, output: torch.Tensor) -> dict:
        fields = {'U': output[:, :3], 'p': output[:, 3:4], 'k': output[:, 4:5], 'epsilon': output[:, 5:6],
                  'nut': output[:, 6:7]}
        if output.shape[1] > 7:
            fields['residual'] = output[:, 7:8]
        return fields
