"""Fused pieces of the reference's training step around the hot path (csrc/train_glue.cu):

* `WeightedMSELoss` — same constructor and forward(pred, target, pressure_ref_weight) as the reference's criterion
  (/root/reference/normalization.py:136-250, built at train.py:352): field-wise weighted MSE over [U(3), p, k, epsilon, nut]
  with the pressure-mean anchor.  CUDA inputs with use_fieldwise=True run as 2 + 1 kernels (forward reduction, gradient)
  instead of ~25 + ~25 torch kernels; anything else takes the reference's formula in torch.
* `FusedClipAdam` — clip_grad_norm_(max_norm) + torch.optim.Adam(lr, betas, eps, weight_decay).step() (train.py:188-189,
  369) over ONE flat parameter / gradient buffer: 3 kernels per step, no host synchronisation, capturable in a CUDA graph.
  The parameters (and their .grad) become views into the flat buffers."""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch

from . import _lib
from .ops import _dt, _p, _stream

_DEFAULT_W = {'U': 1.0, 'p': 3.0, 'k': 0.5, 'epsilon': 0.5, 'nut': 0.5}


class _WMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, fw, prw: float):
        lib = _lib.load()
        n = pred.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        coef = torch.empty(8, dtype=torch.float32, device=pred.device)
        ws = torch.empty(int(lib.b2g_wmse_workspace_bytes()), dtype=torch.uint8, device=pred.device)
        _lib.check(lib.b2g_wmse_fwd(_p(pred), pred.stride(0), _p(target), target.stride(0), n, _dt(pred), _p(fw), float(prw),
                                    _p(loss), _p(coef), _p(ws), _stream()), "wmse_fwd")
        ctx.save_for_backward(pred, target, coef)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        pred, target, coef = ctx.saved_tensors
        n = pred.shape[0]
        g = g.reshape(1).float().contiguous()
        dpred = torch.zeros_like(pred) if pred.shape[1] > 7 else torch.empty_like(pred)
        _lib.check(_lib.load().b2g_wmse_bwd(_p(pred), pred.stride(0), _p(target), target.stride(0), n, _dt(pred), _p(coef),
                                            _p(g), _p(dpred), dpred.stride(0), _stream()), "wmse_bwd")
        return dpred, None, None, None


class WeightedMSELoss(torch.nn.Module):
    def __init__(self, field_weights: Optional[Dict[str, float]] = None, use_fieldwise: bool = True,
                 pressure_ref_weight: float = 0.1):
        super().__init__()
        self.field_weights = dict(_DEFAULT_W) if field_weights is None else field_weights
        self.use_fieldwise = use_fieldwise
        self.pressure_ref_weight = pressure_ref_weight
        g = self.field_weights.get
        self.weights = torch.tensor([g('U', 1.0)] * 3 + [g('p', 1.0), g('k', 0.5), g('epsilon', 0.5), g('nut', 0.5)])
        self._fw = {}

    def forward(self, pred: torch.Tensor, target: torch.Tensor, pressure_ref_weight: float = 0.1) -> torch.Tensor:
        if (self.use_fieldwise and pred.is_cuda and pred.dim() == 2 and pred.shape[1] >= 7 and pred.shape == target.shape
                and pred.dtype == target.dtype and pred.dtype in (torch.float32, torch.bfloat16) and pred.stride(1) == 1
                and target.stride(1) == 1 and pred.shape[0] > 0 and not target.requires_grad):
            fw = self._fw.get(pred.device)
            if fw is None:
                g = self.field_weights.get
                fw = self._fw[pred.device] = torch.tensor([g('U', 1.0), g('p', 1.0), g('k', 0.5), g('epsilon', 0.5), g('nut', 0.5)],
                                                           dtype=torch.float32, device=pred.device)
            return _WMSEFn.apply(pred, target, fw, pressure_ref_weight)
        return self._reference_formula(pred, target, pressure_ref_weight)

    def _reference_formula(self, pred, target, prw):          # normalization.py:188-250
        if self.use_fieldwise:
            g = self.field_weights.get
            mse = lambda a, b: torch.mean((pred[:, a:b] - target[:, a:b]) ** 2)
            p_loss = mse(3, 4)
            if prw > 0:
                p_loss = p_loss + prw * (torch.mean(pred[:, 3:4]) - torch.mean(target[:, 3:4])) ** 2
            return (g('U', 1.0) * mse(0, 3) + g('p', 1.0) * p_loss + g('k', 0.5) * mse(4, 5) + g('epsilon', 0.5) * mse(5, 6)
                    + g('nut', 0.5) * mse(6, 7))
        w = self.weights.to(pred.device)
        return (((pred - target) ** 2) * w.unsqueeze(0)).mean()


class FusedClipAdam:
    """`opt = FusedClipAdam(model.parameters(), lr=3e-4, weight_decay=1e-5, max_grad_norm=1.0)`;
    `opt.zero_grad(); loss.backward(); opt.step()` == zero_grad + clip_grad_norm_(params, max_grad_norm) + Adam.step()."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedClipAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FusedClipAdam: fp32 parameters on one CUDA device are required (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)      # (steps, last gradient norm)
        lib = _lib.load()
        self._ws = torch.empty(int(lib.b2g_adam_workspace_bytes()), dtype=torch.uint8, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)            # the parameter now lives in the flat buffer
                p.grad = self.grad[off:off + k].view_as(p)            # autograd accumulates into the flat gradient buffer
                off += k
        self.param_groups = [{"params": self.params, "lr": lr}]

    def zero_grad(self, set_to_none: bool = False):
        self.grad.zero_()                                                 # one memset; .grad views stay in place
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
                    p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * 4:
                off = (p.data_ptr() - self.flat.data_ptr()) // 4
                p.grad = self.grad[off:off + p.numel()].view_as(p)

    def step(self):
        lr = self.param_groups[0]["lr"]
        _lib.check(_lib.load().b2g_clip_adam_step(_p(self.flat), _p(self.grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                                  self.flat.numel(), float(self.max_grad_norm), float(lr), float(self.betas[0]),
                                                  float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                                  _p(self.state), _p(self._ws), _stream()), "clip_adam_step")

    @property
    def grad_norm(self) -> torch.Tensor:
        return self.state[1]
