"""CPU-only, runs only where /root/reference exists (this container, not the GPU box): the reference's
own gnn_model.py / train.py import and construct on top of the drop-in modules, and the FlowGNN mirror
used on the GPU box has the reference's exact state_dict layout."""
import importlib
import os
import sys

import pytest
import torch

REF = os.environ.get("B2G_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gnn_model.py")), reason="reference not mounted")


@pytest.fixture()
def ref_modules():
    from gnn_bfs_rans_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data",
                                             "graph_constructor", "gnn_model")}
    dropin.install()
    sys.path.insert(0, REF)
    sys.modules.pop("gnn_model", None)
    try:
        yield importlib.import_module("gnn_model")
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("lt", ["GCN", "GAT", "GIN", "Transformer"])
def test_reference_flowgnn_builds_on_dropin_and_matches_mirror(ref_modules, lt):
    from gnn_bfs_rans_b200.flow_model import FlowGNN as Mirror
    torch.manual_seed(0)
    ref = ref_modules.FlowGNN(input_dim=3, hidden_dim=32, output_dim=7, num_layers=2, layer_type=lt, dropout=0.1)
    torch.manual_seed(0)
    mir = Mirror(input_dim=3, hidden_dim=32, output_dim=7, num_layers=2, layer_type=lt, dropout=0.1)
    sr, sm = ref.state_dict(), mir.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    for k in sr:
        assert sr[k].shape == sm[k].shape and torch.equal(sr[k], sm[k]), k     # same init order under the same seed
    assert type(ref.gnn_layers[0]).__module__.endswith("nn")                    # our layer classes, not PyG's
    mir.load_state_dict(sr)                                                     # checkpoints are interchangeable
    # the reference's RuntimeError re-wrap (gnn_model.py:173-181) sees our "no CPU fallback" RuntimeError
    with pytest.raises(RuntimeError, match="Message passing failed"):
        ref(torch.randn(5, 3), torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(ValueError):
        ref_modules.FlowGNN(layer_type="SAGE")


def test_reference_train_module_imports_on_dropin(ref_modules):
    sys.path.insert(0, REF)
    try:
        for mod in ("normalization", "openfoam_loader"):
            importlib.import_module(mod)
        sys.modules.pop("train", None)
        train = importlib.import_module("train")           # imports tqdm, Data/Batch, GraphConstructor, FlowGNN
        from gnn_bfs_rans_b200 import Batch, Data
        b = train.collate_fn([Data(x=torch.zeros(3, 3), edge_index=torch.tensor([[0, 1], [1, 2]]), num_nodes=3)])
        assert isinstance(b, Batch) and b.batch.tolist() == [0, 0, 0]
        assert train.GraphConstructor.__module__.endswith("graph_constructor")
    finally:
        sys.path.remove(REF)
        sys.modules.pop("train", None)
