"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Sampled-row parity at sizes where the full fp64 oracle cannot run (cfg3 / cfg4: the [E, F] / [E, H, C] fp64
temporaries of oracle/layers_oracle.py would be 143 GB / 571 GB at cfg4).  For a set S of target rows the four layers
only look one hop back: out_i depends on x_j of the sources j of i, and — GCNConv — on the in-degree of i and of every j
(/root/reference/gnn_model.py:63-80,166-170 via PyG; oracle/layers_oracle.py gcn_norm / segment_softmax).  So the layer
outputs at S of the sub-problem

    R1    = S  ∪  sources of the edges into S
    edges = ALL edges whose target is in R1            (so every node of R1 keeps its full in-degree)
    nodes = R1 ∪ sources of those edges

equal the outputs at S on the whole graph, exactly, for all four layer types; `closure_subgraph` cuts that sub-problem out
(edge order preserved) and `layer_rows` runs oracle/layers_oracle.py on it in fp64.  bench.py calls this OUTSIDE its timed
regions as the checker of the timed cfg3 / cfg4 outputs (`parity_check`), tests/ use it at mid sizes."""
from __future__ import annotations

import torch

from . import layers_oracle as lo


def closure_subgraph(edge_index: torch.Tensor, targets: torch.Tensor, num_nodes: int):
    """-> (nodes int64 [n_sub] ascending global ids, ei_sub int64 [2, e_sub] in sub-numbering and original edge order,
    pos int64 [len(targets)] = row of each target in the sub-numbering).  Runs on edge_index's device (torch index ops)."""
    dev = edge_index.device
    src, dst = edge_index[0], edge_index[1]
    targets = targets.to(dev).long()
    in_s = torch.zeros(num_nodes, dtype=torch.bool, device=dev)
    in_s[targets] = True
    r1 = in_s.clone()
    r1[src[in_s[dst]]] = True
    keep = r1[dst]
    s2, d2 = src[keep], dst[keep]
    in_n = r1.clone()
    in_n[s2] = True
    nodes = torch.nonzero(in_n).squeeze(1)
    g2l = torch.full((num_nodes,), -1, dtype=torch.int64, device=dev)
    g2l[nodes] = torch.arange(nodes.numel(), device=dev)
    return nodes, torch.stack([g2l[s2], g2l[d2]]), g2l[targets]


def layer_rows(kind: str, params: dict, x_sub: torch.Tensor, ei_sub: torch.Tensor, pos: torch.Tensor, heads: int = 4,
               edge_attr_sub=None):
    """fp64 oracle rows out[pos] of layer `kind` ('GCN' | 'GAT' | 'GIN' | 'Transformer') on the sub-problem.
    params: the layer's state_dict (any dtype / device); x_sub: the feature rows of `nodes` AS THE KERNEL SAW THEM
    (bf16 values are exact in fp64)."""
    p = {k: v.detach().double().cpu() for k, v in params.items()}
    x = x_sub.detach().double().cpu()
    ei = ei_sub.cpu()
    with torch.no_grad():
        if kind == "GCN":
            out = lo.gcn_conv(x, ei, p["lin.weight"], p.get("bias"))
        elif kind == "GAT":
            out = lo.gat_conv(x, ei, p["lin.weight"], p["att_src"], p["att_dst"], p.get("bias"), heads=heads)
        elif kind == "GIN":
            out = lo.gin_conv(x, ei, lo.gin_mlp(p["nn.0.weight"], p["nn.0.bias"], p["nn.2.weight"], p["nn.2.bias"]),
                              eps=float(p["eps"]) if "eps" in p else 0.0)
        elif kind == "Transformer":
            out = lo.transformer_conv(x, ei, p["lin_query.weight"], p["lin_query.bias"], p["lin_key.weight"],
                                      p["lin_key.bias"], p["lin_value.weight"], p["lin_value.bias"],
                                      p["lin_skip.weight"], p["lin_skip.bias"], heads=heads,
                                      edge_attr=edge_attr_sub, we=p.get("lin_edge.weight"))
        else:
            raise ValueError(kind)
    return out[pos.cpu()]


def pick_rows(n: int, count: int, plane: int = 0, seed: int = 0) -> torch.Tensor:
    """`count` distinct target rows of an n-row problem: the first and last rows (first / last panel of the row schedule and,
    on the lexicographic hex block, the z = 0 and z = max boundary planes), rows around multiples of `plane` (the other
    faces of the block / slab boundaries) and uniformly random interior rows."""
    g = torch.Generator().manual_seed(seed)
    q = max(count // 8, 1)
    parts = [torch.arange(0, min(q, n)), torch.arange(max(n - q, 0), n)]
    if plane > 0 and n > 2 * plane:
        k = torch.randint(1, n // plane, (q,), generator=g) * plane
        parts += [k, (k - 1).clamp_min(0), (k + 1).clamp_max(n - 1)]
    rows = torch.unique(torch.cat(parts))
    if rows.numel() > count:                       # evenly thinned, first and last row kept
        rows = rows[torch.linspace(0, rows.numel() - 1, count).round().long()]
    while rows.numel() < min(count, n):
        extra = torch.randint(0, n, (2 * (count - rows.numel()) + 16,), generator=g)
        new = torch.unique(torch.cat([rows, extra]))
        if new.numel() > count:                    # drop random extras only: the deterministic rows all stay
            ex = new[~torch.isin(new, rows)]
            new = torch.cat([rows, ex[torch.randperm(ex.numel(), generator=g)[:count - rows.numel()]]]).sort().values
        rows = new
    return rows
