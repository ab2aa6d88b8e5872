"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Pure-torch CPU restatement of the PyG (torch-geometric >= 2.3, unpinned in
/root/reference/requirements.txt:2, NOT installed here, no source on disk) message-passing layers
the reference calls at /root/reference/gnn_model.py:62-80 (ctors) and :165-172 (forward).

PARITY UNPINNED: the reference has no test or golden vector that pins any layer output and PyG
cannot be imported in this container, so this file restates PyG's published algorithm
(torch_geometric/nn/conv/{gcn,gat,gin,transformer}_conv.py, torch_geometric/utils/{softmax,loop}.py)
op-for-op with the same torch CPU kernels PyG dispatches to without torch_scatter
(index_select / scatter_add_ / scatter_reduce_(amax) / F.linear).  Run in float64 it is the truth
for the parity gates (1e-5 fp32 / 2e-2 bf16, relative to the tensor's max-abs); run in float32 it
is the timed "PyG CPU path" of bench.py's cpu_baseline.

flow = source_to_target: j = edge_index[0] (source), i = edge_index[1] (target, aggregation index).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ---- torch_geometric.utils.loop ------------------------------------------------------------------
def remove_self_loops(edge_index):
    return edge_index[:, edge_index[0] != edge_index[1]]


def add_self_loops(edge_index, num_nodes):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)   # appended at the END


def replace_self_loops(edge_index, num_nodes):
    """GATConv: remove_self_loops + add_self_loops; GCN's add_remaining_self_loops gives the same
    edge list for unweighted graphs."""
    return add_self_loops(remove_self_loops(edge_index), num_nodes)


# ---- torch_geometric.utils.scatter / softmax -------------------------------------------------------
def scatter_sum(src, index, num_nodes):
    out = src.new_zeros((num_nodes,) + tuple(src.shape[1:]))
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return out.scatter_add_(0, idx, src)


def segment_softmax(src, index, num_nodes):
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    m = src.new_zeros((num_nodes,) + tuple(src.shape[1:]))
    m = m.scatter_reduce_(0, idx, src.detach(), reduce='amax', include_self=False)
    out = (src - m.index_select(0, index)).exp()
    z = scatter_sum(out, index, num_nodes) + 1e-16
    return out / z.index_select(0, index)


# ---- layers --------------------------------------------------------------------------------------
def gcn_norm(edge_index, num_nodes, dtype):
    ei = replace_self_loops(edge_index, num_nodes)
    row, col = ei[0], ei[1]
    w = torch.ones(ei.shape[1], dtype=dtype, device=ei.device)
    deg = scatter_sum(w, col, num_nodes)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float('inf'), 0)
    return ei, dis[row] * w * dis[col]


def gcn_conv(x, edge_index, weight, bias=None):
    """GCNConv(in,out).forward(x, edge_index): lin (no bias) -> normalised scatter-sum -> + bias."""
    N = x.shape[0]
    ei, w = gcn_norm(edge_index, N, x.dtype)
    h = F.linear(x, weight)
    out = scatter_sum(w.view(-1, 1) * h.index_select(0, ei[0]), ei[1], N)
    return out if bias is None else out + bias


def gat_conv(x, edge_index, weight, att_src, att_dst, bias=None, heads=4, concat=False,
             negative_slope=0.2, dropout=0.0, training=False, return_alpha=False,
             edge_attr=None, we=None, att_edge=None):
    """GATConv(in,out,heads,concat,dropout).forward(x, edge_index[, edge_attr]) with add_self_loops=True.
    edge_dim (PyG GATConv.forward / edge_updater): with `we` = lin_edge.weight [H*C, edge_dim] and `att_edge` [1,H,C],
    the self loops are dropped WITH their attributes, every node's new loop gets the mean attribute of its incoming edges
    (fill_value='mean'; 0 for a node without one), and alpha_edge = (lin_edge(edge_attr).view(-1,H,C) * att_edge).sum(-1)
    joins a_src[j] + a_dst[i] in front of the LeakyReLU.  The messages stay x_j W (no edge term).
    (the product: nn.GATConv(edge_dim=4) -> functional.GATZFn with ve / ea.)"""
    N = x.shape[0]
    H = heads
    C = weight.shape[0] // H
    xs = F.linear(x, weight).view(N, H, C)
    a_s = (xs * att_src.view(1, H, C)).sum(-1)
    a_d = (xs * att_dst.view(1, H, C)).sum(-1)
    ei = replace_self_loops(edge_index, N)
    row, col = ei[0], ei[1]
    logit = a_s.index_select(0, row) + a_d.index_select(0, col)
    if edge_attr is not None and we is not None:
        keep = edge_index[0] != edge_index[1]                               # remove_self_loops(edge_index, edge_attr)
        ea = edge_attr[keep].to(x.dtype)
        tgt = edge_index[1][keep]
        cnt = scatter_sum(torch.ones_like(tgt, dtype=x.dtype), tgt, N).clamp_min(1)
        loop_attr = scatter_sum(ea, tgt, N) / cnt.unsqueeze(-1)            # add_self_loops(..., fill_value='mean')
        e = F.linear(torch.cat([ea, loop_attr], dim=0), we).view(-1, H, C)
        logit = logit + (e * att_edge.view(1, H, C)).sum(-1)
    alpha = F.leaky_relu(logit, negative_slope)
    alpha = segment_softmax(alpha, col, N)
    alpha = F.dropout(alpha, p=dropout, training=training)
    out = scatter_sum(alpha.unsqueeze(-1) * xs.index_select(0, row), col, N)
    out = out.reshape(N, H * C) if concat else out.mean(dim=1)
    if bias is not None:
        out = out + bias
    return (out, ei, alpha) if return_alpha else out


def gin_conv(x, edge_index, mlp, eps=0.0):
    """GINConv(nn).forward(x, edge_index): nn( sum_j x_j + (1+eps) x_i ) on the RAW edge list."""
    N = x.shape[0]
    out = scatter_sum(x.index_select(0, edge_index[0]), edge_index[1], N)
    out = out + (1 + eps) * x
    return mlp(out)


def gin_mlp(w1, b1, w2, b2):
    return lambda h: F.linear(F.relu(F.linear(h, w1, b1)), w2, b2)


def transformer_conv(x, edge_index, wq, bq, wk, bk, wv, bv, ws, bs, heads=4, concat=False,
                     dropout=0.0, training=False, return_alpha=False, edge_attr=None, we=None):
    """TransformerConv(in,out,heads,concat,dropout).forward(x, edge_index[, edge_attr]) (beta=False,
    root_weight=True) on the RAW edge list.  edge_dim (PyG TransformerConv.message): with `we` = lin_edge.weight
    [H*C, edge_dim] (no bias) and an edge_attr, e = lin_edge(edge_attr).view(-1, H, C) is added to key_j before the
    logits and to value_j before the weighted sum."""
    N = x.shape[0]
    H = heads
    C = wq.shape[0] // H
    q = F.linear(x, wq, bq).view(N, H, C)
    k = F.linear(x, wk, bk).view(N, H, C)
    v = F.linear(x, wv, bv).view(N, H, C)
    row, col = edge_index[0], edge_index[1]
    k_j, v_j = k.index_select(0, row), v.index_select(0, row)
    if we is not None and edge_attr is not None:
        e = F.linear(edge_attr.to(x.dtype), we).view(-1, H, C)
        k_j = k_j + e
        v_j = v_j + e
    alpha = (q.index_select(0, col) * k_j).sum(-1) / math.sqrt(C)
    alpha = segment_softmax(alpha, col, N)
    alpha = F.dropout(alpha, p=dropout, training=training)
    out = scatter_sum(v_j * alpha.view(-1, H, 1), col, N)
    out = out.reshape(N, H * C) if concat else out.mean(dim=1)
    out = out + F.linear(x, ws, bs)
    return (out, alpha) if return_alpha else out


def batch_norm(x, weight, bias, running_mean=None, running_var=None, training=True, momentum=0.1, eps=1e-5):
    """torch_geometric.nn.BatchNorm(F).forward == torch.nn.BatchNorm1d(F) (gnn_model.py:87,188)."""
    return F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


# ---- parameter initialisers (torch_geometric.nn.inits) ---------------------------------------------
def glorot_(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


def flow_gnn_forward(x, edge_index, params, layer_type, training=False, edge_attr=None):
    """FlowGNN.forward (gnn_model.py:159-197) with dropout p=0 and BatchNorm in the given mode,
    expressed over the functional oracle layers.  `params` mirrors FlowGNN.state_dict()."""
    p = params
    h = F.linear(x, p['input_proj.weight'], p['input_proj.bias'])
    L = 0
    while f'gnn_layers.{L}.bias' in p or f'gnn_layers.{L}.eps' in p or f'gnn_layers.{L}.lin_skip.weight' in p:
        L += 1
    for i in range(L):
        g = f'gnn_layers.{i}.'
        if layer_type == 'GCN':
            hn = gcn_conv(h, edge_index, p[g + 'lin.weight'], p[g + 'bias'])
        elif layer_type == 'GAT':
            hn = gat_conv(h, edge_index, p[g + 'lin.weight'], p[g + 'att_src'], p[g + 'att_dst'], p[g + 'bias'],
                          edge_attr=edge_attr if (g + 'lin_edge.weight') in p else None,
                          we=p.get(g + 'lin_edge.weight'), att_edge=p.get(g + 'att_edge'))
        elif layer_type == 'GIN':
            hn = gin_conv(h, edge_index, gin_mlp(p[g + 'nn.0.weight'], p[g + 'nn.0.bias'],
                                                 p[g + 'nn.2.weight'], p[g + 'nn.2.bias']))
        elif layer_type == 'Transformer':
            hn = transformer_conv(h, edge_index,
                                  p[g + 'lin_query.weight'], p[g + 'lin_query.bias'],
                                  p[g + 'lin_key.weight'], p[g + 'lin_key.bias'],
                                  p[g + 'lin_value.weight'], p[g + 'lin_value.bias'],
                                  p[g + 'lin_skip.weight'], p[g + 'lin_skip.bias'],
                                  edge_attr=edge_attr, we=p.get(g + 'lin_edge.weight'))   # gnn_model.py:170 passes edge_attr
        else:
            raise ValueError(layer_type)
        h = h + hn
        b = f'batch_norms.{i}.module.'
        if b + 'weight' in p:
            h = F.batch_norm(h, None if training else p[b + 'running_mean'],
                             None if training else p[b + 'running_var'],
                             p[b + 'weight'], p[b + 'bias'], training, 0.1, 1e-5)
        h = F.relu(h)
    for j in (0, 3, 6):
        h = F.relu(F.linear(h, p[f'output_proj.{j}.weight'], p[f'output_proj.{j}.bias']))
    return F.linear(h, p['output_proj.8.weight'], p['output_proj.8.bias'])
