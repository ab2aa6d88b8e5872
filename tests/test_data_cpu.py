"""CPU: drop-in Data / Batch semantics (torch_geometric.data surface the reference uses: graph_constructor.py:262-267,
train.py:111-137,155,167), incl. the deferred collation of Batch.from_data_list (host materialisation path)."""
import torch

from gnn_bfs_rans_b200.data import Batch, Data


def _samples():
    g = torch.Generator().manual_seed(0)
    out = []
    for n, e in ((5, 8), (3, 2), (7, 12)):
        out.append(Data(x=torch.randn(n, 3, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g),
                        edge_attr=torch.randn(e, 4, generator=g), y=torch.randn(n, 7, generator=g), num_nodes=n))
    return out


def test_deferred_batch_materialises_to_pyg_semantics_on_the_host():
    ds = _samples()
    b = Batch.from_data_list(ds)
    assert '_pending' in b.__dict__                      # nothing concatenated yet
    assert b.num_graphs == 3
    ei = torch.cat([ds[0].edge_index, ds[1].edge_index + 5, ds[2].edge_index + 8], dim=1)
    assert torch.equal(b.edge_index, ei) and '_pending' not in b.__dict__
    assert torch.equal(b.x, torch.cat([d.x for d in ds])) and torch.equal(b.y, torch.cat([d.y for d in ds]))
    assert torch.equal(b.edge_attr, torch.cat([d.edge_attr for d in ds]))
    assert b.batch.tolist() == [0] * 5 + [1] * 3 + [2] * 7 and b.ptr.tolist() == [0, 5, 8, 15] and b.num_nodes == 15
    b2 = Batch.from_data_list(ds).to('cpu')              # .to(cpu) = host path
    assert torch.equal(b2.edge_index, ei)
    b3 = Batch.from_data_list(ds)
    b3.x = torch.zeros(15, 3)                             # a mutation first materialises, then applies
    assert torch.equal(b3.x, torch.zeros(15, 3)) and torch.equal(b3.edge_index, ei)


def test_single_graph_batch_and_attribute_bag():
    d = _samples()[0]
    b = Batch.from_data_list([d])
    assert torch.equal(b.edge_index, d.edge_index) and b.batch.tolist() == [0] * 5
    d.x = torch.ones(5, 3)                                # the reference mutates graph.x / y / num_nodes (train.py:111-137)
    d.num_nodes = 5
    assert d.num_nodes == 5 and 'x' in d and d.keys()[0] == 'x'
    assert Batch.from_data_list([]).num_graphs == 1
