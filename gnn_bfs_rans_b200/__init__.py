"""Importable alias of the product package, which lives in the directory `gnn-bfs-rans_b200/`
(a hyphenated name cannot be imported directly).  `import gnn_bfs_rans_b200` executes that
directory's __init__.py as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gnn-bfs-rans_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
