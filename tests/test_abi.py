"""CPU-only: libb2g.so loads and exports every symbol include/b2g.h declares, with the argument
counts the ctypes table binds; the host-side mirror imports; no compute calls."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    src = open(os.path.join(ROOT, "include", "b2g.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|void|const char\*)\s+(b2g_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_exports_every_declared_symbol():
    from gnn_bfs_rans_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()
    decls = _header_decls()
    assert len(decls) >= 30
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} declared in include/b2g.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes binding"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(decls)
    assert lib.b2g_version() == 100
    assert lib.b2g_error_string(-2).decode().startswith("b2g:")
    # argument validation is host-side C: callable without a GPU
    assert lib.b2g_seg_sum(None, 0, None, 0, None, 0, -1, 8, 0, None, None, None, None, 0.0, None, 0, None) == -1
    assert lib.b2g_csr_workspace_bytes(10, 5) > 0
    assert lib.b2g_linear_impl(1000, 256, 256, 1, 0) in (1, 2)


def test_no_fallback_without_gpu_or_library():
    import torch
    import gnn_bfs_rans_b200 as b2g
    layer = b2g.nn.GCNConv(8, 8)
    with pytest.raises(RuntimeError):
        layer(torch.randn(4, 8), torch.tensor([[0, 1], [1, 0]]))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            b2g.GraphConstructor(dict(owner=[0], neighbour=[], cell_centers=[[0, 0, 0]], n_cells=1)).build_graph()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gnn-bfs-rans_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_module_surface_and_state_dict_keys():
    import torch
    import gnn_bfs_rans_b200 as b2g
    from gnn_bfs_rans_b200.flow_model import FlowGNN
    m = FlowGNN(3, 32, 7, 2, 'GAT')
    keys = set(m.state_dict().keys())
    for k in ("input_proj.weight", "gnn_layers.0.lin.weight", "gnn_layers.0.att_src", "gnn_layers.1.bias",
              "batch_norms.0.module.running_mean", "output_proj.8.bias"):
        assert k in keys, k
    assert m.gnn_layers[0].att_src.shape == (1, 4, 32) and m.gnn_layers[0].lin.weight.shape == (128, 32)
    t = FlowGNN(3, 32, 7, 1, 'Transformer').state_dict()
    assert t["gnn_layers.0.lin_skip.weight"].shape == (32, 32) and t["gnn_layers.0.lin_key.bias"].shape == (128,)
    gi = FlowGNN(3, 32, 7, 1, 'GIN').state_dict()
    assert "gnn_layers.0.eps" in gi and gi["gnn_layers.0.nn.2.weight"].shape == (32, 32)
    with pytest.raises(ValueError):
        FlowGNN(layer_type='SAGE')
    # glorot bound of PyG's inits
    torch.manual_seed(0)
    w = b2g.nn.GCNConv(64, 64).lin.weight
    assert float(w.abs().max()) <= (6 / 128) ** 0.5 + 1e-6
    assert float(b2g.nn.GCNConv(64, 64).bias.abs().max()) == 0.0


def test_data_and_batch_semantics():
    import torch
    from gnn_bfs_rans_b200 import Data, Batch
    d = Data(x=torch.zeros(4, 3), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.zeros(2, 4), num_nodes=4)
    d.y = torch.ones(4, 7)
    d.num_nodes = 3
    assert d.num_nodes == 3 and d.y.shape == (4, 7) and d.batch is None
    b = Batch.from_data_list([d])
    assert b.batch.tolist() == [0, 0, 0] and b.edge_index.tolist() == [[0, 1], [1, 0]] and b.num_graphs == 1
    b3 = Batch.from_data_list([d, d.clone(), d.clone()])
    assert b3.edge_index[:, -2:].tolist() == [[6, 7], [7, 6]] and b3.x.shape == (12, 3) and b3.num_nodes == 9


def test_identity_cache_follows_the_tensor_object_and_its_version():
    """graph.identity_cached (used for edge_attr in CSR order): a hit needs the same tensor object, unmodified."""
    from gnn_bfs_rans_b200.graph import identity_cached
    store, calls = {}, []
    a = torch.arange(6.0)

    def make(t):
        calls.append(1)
        return t * 2

    r1 = identity_cached(store, "k", a, lambda: make(a))
    r2 = identity_cached(store, "k", a, lambda: make(a))
    assert r1 is r2 and len(calls) == 1
    a.add_(1)                                           # in-place change bumps the version counter
    r3 = identity_cached(store, "k", a, lambda: make(a))
    assert len(calls) == 2 and torch.equal(r3, a * 2)
    b = a.clone()                                       # equal content, different object
    identity_cached(store, "k", b, lambda: make(b))
    assert len(calls) == 3
    del b                                               # a dead referent never matches a new tensor
    c = torch.arange(6.0)
    identity_cached(store, "k", c, lambda: make(c))
    assert len(calls) == 4
