"""Makes the reference's scripts run byte-for-byte unchanged on this implementation.

`install()` registers this package under the import names the reference uses
(train.py:9,18; gnn_model.py:8-10; graph_constructor.py:8):
    torch_geometric, torch_geometric.nn, torch_geometric.data, torch_geometric.utils  -> nn.py / data.py
    graph_constructor                                                                -> graph_constructor.py
`python -m gnn_bfs_rans_b200.dropin /path/to/reference/train.py --epochs 2 ...` then executes the
script with runpy, with the reference directory on sys.path for its other modules
(gnn_model, openfoam_loader, normalization)."""
from __future__ import annotations

import os
import runpy
import sys
import types


def install(force: bool = True, graph_constructor: bool = True):
    from . import data as _data
    from . import graph_constructor as _gc
    from . import nn as _nn

    if not force:
        try:
            import torch_geometric  # noqa: F401  (a real PyG wins unless forced)
            return False
        except ImportError:
            pass
    pkg = types.ModuleType("torch_geometric")
    pkg.__path__ = []  # mark as package
    pkg.__version__ = "2.3.0+b2g"
    mnn = types.ModuleType("torch_geometric.nn")
    for name in ("MessagePassing", "global_mean_pool", "GCNConv", "GATConv", "GINConv", "TransformerConv",
                 "BatchNorm", "Linear"):
        setattr(mnn, name, getattr(_nn, name))
    mdata = types.ModuleType("torch_geometric.data")
    mdata.Data, mdata.Batch = _data.Data, _data.Batch
    pkg.nn, pkg.data = mnn, mdata
    sys.modules["torch_geometric"] = pkg
    sys.modules["torch_geometric.nn"] = mnn
    sys.modules["torch_geometric.data"] = mdata
    if graph_constructor:
        mgc = types.ModuleType("graph_constructor")
        mgc.GraphConstructor = _gc.GraphConstructor
        mgc.__file__ = _gc.__file__
        sys.modules["graph_constructor"] = mgc
    return True


def run(script: str, argv=None):
    script = os.path.abspath(script)
    install()
    ref_dir = os.path.dirname(script)
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    old_argv = sys.argv
    sys.argv = [script] + list(argv or [])
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("usage: python -m gnn_bfs_rans_b200.dropin <reference script.py> [script args...]")
        sys.exit(2)
    run(sys.argv[1], sys.argv[2:])
