"""Fused optimizers over a flat parameter arena.

The reference trains with `torch.optim.SGD(lr, momentum=0.9, weight_decay)` + a LambdaLR warm-up
schedule (train_baseline_cv.py:111-119, :180-182) and the ensemble with `Adam(lr=1e-4)`
(train_ensemble_whole_dataset.py:53).  Stock optimizers launch several kernels per parameter tensor
(~100 tensors per ViT), which dominates a B=4 step.  Here every parameter (and its gradient) is a view
into ONE contiguous fp32 buffer and a step is a single `vit3d_sgd_step` / `vit3d_adam_step` launch.
The classes subclass `torch.optim.Optimizer`, so `utils/scheduler.py`'s LambdaLR schedules and the
scripts' `optimizer.step(); scheduler.step(); optimizer.zero_grad()` sequence work unchanged.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from ._lib import call, ptr, stream
from .functional import enable_direct_grads, invalidate_weight_shadows


class FlatArena:
    """Re-homes parameters (and their gradients) as views into two flat fp32 buffers.  Build it AFTER the
    model has been moved to its device."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("all parameters must be fp32 on one device")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.numel, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.grad_views = []
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                view = self.flat[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                gv = self.flat_grad[off:off + n].view_as(p)
                if p.grad is not None:
                    gv.copy_(p.grad)
                p.grad = gv
                self.grad_views.append(gv)
                off += n

    def sync_grads(self):
        """Make sure every p.grad IS its arena view (a `zero_grad(set_to_none=True)` or a fresh autograd
        tensor may have replaced it); copies stray gradients in."""
        for p, gv in zip(self.params, self.grad_views):
            g = p.grad
            if g is gv:
                continue
            if g is None:
                gv.zero_()
            elif g.data_ptr() != gv.data_ptr():
                gv.copy_(g)
            p.grad = gv

    def zero_grad(self):
        if self.flat_grad.is_cuda:      # a memset node, not a framework fill kernel (CUDA-graph'd steps: kernels are all ours)
            call("vit3d_memset_zero", ptr(self.flat_grad), self.flat_grad.numel() * 4, stream())
        else:
            self.flat_grad.zero_()
        for p, gv in zip(self.params, self.grad_views):
            p.grad = gv


class _FlatOptimizer(torch.optim.Optimizer):
    def __init__(self, params, defaults):
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("fused flat optimizers take one parameter group")
        super().__init__(params, defaults)
        self.arena = FlatArena(self.param_groups[0]["params"])
        enable_direct_grads(True)      # backward kernels add straight into the arena views
        self._steps = 0
        # device-resident lr / step for CUDA-graph replays (see graphs.GraphedTrainStep); None = host values
        self.lr_dev = None
        self.step_dev = None

    def zero_grad(self, set_to_none: bool = False):
        self.arena.zero_grad()


class FusedSGD(_FlatOptimizer):
    """torch.optim.SGD(momentum, weight_decay, dampening=0, nesterov=False) semantics in one launch."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.momentum_buffer = torch.zeros_like(self.arena.flat) if momentum != 0.0 else None

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        a = self.arena
        a.sync_grads()
        call("vit3d_sgd_step", ptr(a.flat), ptr(a.flat_grad), ptr(self.momentum_buffer), a.numel, float(g["lr"]),
             float(g["momentum"]), float(g["weight_decay"]), 0, float(grad_scale), ptr(self.lr_dev), stream())
        self._steps += 1
        invalidate_weight_shadows()   # the kernel updated the weights behind torch's version counters
        return loss


class FusedAdam(_FlatOptimizer):
    """torch.optim.Adam (no amsgrad) semantics in one launch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(self.arena.flat)
        self.exp_avg_sq = torch.zeros_like(self.arena.flat)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        a = self.arena
        a.sync_grads()
        self._steps += 1
        call("vit3d_adam_step", ptr(a.flat), ptr(a.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), a.numel,
             float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
             self._steps, float(grad_scale), ptr(self.lr_dev), ptr(self.step_dev), stream())
        invalidate_weight_shadows()
        return loss
