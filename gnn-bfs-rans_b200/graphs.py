"""CUDA-graph capture of a forward pass on a static mesh.

The shipped BFS case has 12 k nodes: a 4-layer FlowGNN forward is ~100 kernel launches of a few microseconds each, so the
step is bound by launch latency, not by any kernel.  The mesh (edge_index) of the reference is the same for every sample
(`train.py:120-133` builds one graph per time directory of ONE case), so the whole forward can be captured once and
replayed: inputs are copied into a static buffer, one `cudaGraphLaunch` runs every kernel back to back.

Capture needs a forward without host synchronisation: `FlowGNN(validate_edges=False)` (the reference's two `.item()`
checks at gnn_model.py:131-132 are done once, eagerly, before the capture) and eval mode (no dropout seeds)."""
from __future__ import annotations

import torch


class GraphedForward:
    """`out = GraphedForward(model, x_example, edge_index)(x)` == `model(x, edge_index)` (bit-identical: same kernels in
    the same order), replayed from a CUDA graph.  `edge_index` must stay alive and unchanged; `x` must keep its shape."""

    def __init__(self, model: torch.nn.Module, x: torch.Tensor, edge_index: torch.Tensor, warmup: int = 3):
        if not x.is_cuda or not edge_index.is_cuda:
            raise RuntimeError("b2g.graphs: CUDA tensors required (no CPU fallback)")
        if model.training:
            raise RuntimeError("b2g.graphs: capture the eval-mode forward (model.eval()); dropout seeds are host state")
        self.model, self.edge_index = model, edge_index
        n = x.shape[0]
        if edge_index.numel() and (int(edge_index.min()) < 0 or int(edge_index.max()) >= n):   # gnn_model.py:130-149, once
            raise ValueError("edge_index has entries outside [0, num_nodes): filter it before capturing")
        self._restore = getattr(model, "validate_edges", None)
        if self._restore is not None:
            model.validate_edges = False
        self.x = x.clone()
        side = torch.cuda.Stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                       # builds the CSR cache, hints, kernel attributes eagerly
                model(self.x, edge_index)
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(self.x, edge_index)
        if self._restore is not None:
            model.validate_edges = self._restore

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.x.shape or x.dtype != self.x.dtype:
            raise ValueError(f"captured for x {tuple(self.x.shape)} {self.x.dtype}, got {tuple(x.shape)} {x.dtype}")
        self.x.copy_(x)
        self.graph.replay()
        return self.out
