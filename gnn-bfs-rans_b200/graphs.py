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
        # the captured kernels hold raw pointers into this Graph's CSR buffers: own it for the lifetime of the capture
        # (graph_of's LRU would otherwise free it once enough other meshes have been seen)
        from .graph import graph_of
        self._graph = graph_of(edge_index, n)
        self._graph._pinned = True
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


class GraphedTrainStep:
    """One optimisation step — zero_grad, forward, loss, backward, (clip), optimizer.step — on a static mesh, captured
    once and replayed from a CUDA graph (train.py:170-189 on the shipped 12 k-node case is ~400 launches of a few
    microseconds: launch-bound).  `step(x, y)` copies the sample into the static buffers, replays, and returns the loss
    tensor of that step (device; read it with .item() only when needed).

    Dropout (attention dropout inside GATConv / TransformerConv, the fused glue of FlowGNN(fused_glue=True)) draws new
    masks on every replay through the library's device-side epoch (include/b2g.h b2g_dropout_epoch_advance), captured as
    the first node of the graph.  torch's own nn.Dropout uses torch's graph-safe Philox offsets.

    Requirements: `optimizer` must be capturable (e.g. torch.optim.Adam(..., capturable=True)); `loss_fn(out, y)` must not
    synchronise; BatchNorm running statistics, parameters and optimizer state are updated in place as usual."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn, x: torch.Tensor,
                 y: torch.Tensor, edge_index: torch.Tensor, max_grad_norm: float = None, warmup: int = 3):
        from . import _lib
        if not x.is_cuda or not edge_index.is_cuda:
            raise RuntimeError("b2g.graphs: CUDA tensors required (no CPU fallback)")
        for grp in optimizer.param_groups:
            if "capturable" in grp and not grp["capturable"]:
                raise RuntimeError("b2g.graphs: construct the optimizer with capturable=True")
        n = x.shape[0]
        if edge_index.numel() and (int(edge_index.min()) < 0 or int(edge_index.max()) >= n):
            raise ValueError("edge_index has entries outside [0, num_nodes): filter it before capturing")
        self.model, self.optimizer, self.edge_index = model, optimizer, edge_index
        from .graph import graph_of
        self._graph = graph_of(edge_index, n)         # owns the CSR buffers the captured kernels point into
        self._graph._pinned = True
        self._restore = getattr(model, "validate_edges", None)
        if self._restore is not None:
            model.validate_edges = False
        self.x, self.y = x.clone(), y.clone()
        lib = _lib.load()
        params = [p for grp in optimizer.param_groups for p in grp["params"]]

        def one_step():
            _lib.check(lib.b2g_dropout_epoch_advance(torch.cuda.current_stream().cuda_stream), "dropout_epoch_advance")
            optimizer.zero_grad(set_to_none=True)
            loss = loss_fn(model(self.x, edge_index), self.y)
            loss.backward()
            if max_grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
            optimizer.step()
            return loss

        side = torch.cuda.Stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):           # real steps: CSR caches, kernel attributes, optimizer state
                one_step()
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = one_step()
        if self._restore is not None:
            model.validate_edges = self._restore

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.shape != self.x.shape or y.shape != self.y.shape:
            raise ValueError("GraphedTrainStep: x / y shapes differ from the captured ones")
        self.x.copy_(x)
        self.y.copy_(y)
        self.graph.replay()
        return self.loss
