"""TEST INFRASTRUCTURE ONLY. Minimal stand-in for the `torch_geometric` package so that the
reference's graph_constructor.py (which only needs `torch_geometric.data.Data` as an attribute
bag, /root/reference/graph_constructor.py:8,262-267) can be imported unmodified when generating
golden vectors (oracle/make_golden.py). Never imported by the product."""
