"""gnn-bfs-rans_b200 — B200-native (sm_100a) message-passing hot path of Caesar3142/GNN-BFS-RANS.

Host side mirrors the reference's plugin surface:
    nn.{GCNConv, GATConv, GINConv, TransformerConv, BatchNorm, MessagePassing, global_mean_pool}
    data.{Data, Batch}
    graph_constructor.GraphConstructor
    dropin.install() / `python -m gnn_bfs_rans_b200.dropin <reference script>`
    streaming.gcn_forward_host(layer, x_host, edge_index_host): the layer call from host buffers, copies overlapped
    graphs.GraphedForward(model, x, edge_index): the eval forward on a static mesh replayed from one CUDA graph
    training.{WeightedMSELoss, FusedClipAdam}: the reference's criterion and clip + Adam as fused kernels
All arithmetic goes through libb2g.so (C ABI in include/b2g.h, kernels in csrc/)."""
from . import _lib  # noqa: F401
from . import data, dropin, functional, graph, graph_constructor, graphs, mesh, nn, ops, streaming, training  # noqa: F401
from .data import Batch, Data  # noqa: F401
from .graph_constructor import GraphConstructor  # noqa: F401
from .nn import BatchNorm, GATConv, GCNConv, GINConv, TransformerConv  # noqa: F401

__version__ = "0.1.0"
