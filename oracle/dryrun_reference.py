"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Runs one of the reference's own scripts (/root/reference/train.py, inference.py; unchanged) on the CPU with the ORACLE in
place of the kernels: the drop-in module registration (gnn_bfs_rans_b200.dropin: Data / Batch / class names / state_dict
layout) stays, but every layer forward is oracle/layers_oracle.py and the graph builder is oracle/builder_oracle.py.
Purpose: an end-to-end reference for what the drop-in run on the B200 must reproduce — e.g. inference.py on a checkpoint
written by the B200 run gives predictions.npz here, to be compared with the B200's own predictions.npz
(scripts/compare_reference_predictions.py).

    python -m oracle.dryrun_reference /root/reference/inference.py --checkpoint ... --device cpu --output_dir ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def install_oracle_layers():
    import numpy as np
    import torch
    from gnn_bfs_rans_b200 import data as bdata
    from gnn_bfs_rans_b200 import graph_constructor as gcm
    from gnn_bfs_rans_b200 import nn as bnn
    from oracle import builder_oracle as bo
    from oracle import layers_oracle as lo

    def gcn(self, x, ei, edge_weight=None):
        return lo.gcn_conv(x, ei, self.lin.weight, self.bias)

    def gat(self, x, ei, edge_attr=None, size=None, return_attention_weights=None):
        return lo.gat_conv(x, ei, self.lin.weight, self.att_src, self.att_dst, self.bias, heads=self.heads,
                           concat=self.concat, negative_slope=self.negative_slope, dropout=self.dropout,
                           training=self.training)

    def gin(self, x, ei, size=None):
        return lo.gin_conv(x, ei, self.nn, eps=float(self.eps))

    def tconv(self, x, ei, edge_attr=None, return_attention_weights=None):
        return lo.transformer_conv(x, ei, self.lin_query.weight, self.lin_query.bias, self.lin_key.weight,
                                   self.lin_key.bias, self.lin_value.weight, self.lin_value.bias, self.lin_skip.weight,
                                   self.lin_skip.bias, heads=self.heads, concat=self.concat, dropout=self.dropout,
                                   training=self.training)          # edge_dim=None: edge_attr ignored

    bnn.GCNConv.forward, bnn.GATConv.forward, bnn.GINConv.forward, bnn.TransformerConv.forward = gcn, gat, gin, tconv
    bnn.BatchNorm.forward = lambda self, x: self.module(x)
    bnn.Linear.forward = lambda self, x: torch.nn.functional.linear(x, self.weight, self.bias)

    class GraphConstructor:
        def __init__(self, mesh_data):
            self.mesh_data = mesh_data

        def build_graph(self, field_data=None, node_features=None, filter_internal=False, n_internal_cells=None):
            r = bo.build_graph(self.mesh_data, field_data=field_data, node_features=node_features,
                               filter_internal=filter_internal, n_internal_cells=n_internal_cells)
            return bdata.Data(x=torch.from_numpy(np.asarray(r['x'], dtype=np.float32)),
                              edge_index=torch.from_numpy(r['edge_index']), edge_attr=torch.from_numpy(r['edge_attr']),
                              num_nodes=r['num_nodes'])
    gcm.GraphConstructor = GraphConstructor


def main():
    from gnn_bfs_rans_b200 import dropin
    install_oracle_layers()
    script = os.path.abspath(sys.argv[1])
    os.chdir(os.path.dirname(script))
    dropin.run(script, sys.argv[2:])


if __name__ == "__main__":
    main()
