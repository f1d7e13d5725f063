"""FusedAdamW: torch.optim.AdamW semantics (src/train.py:44-49) executed by libvs_b200 kernels.

It is a regular `torch.optim.Optimizer` (same `param_groups` keys: lr, betas, eps, weight_decay),
so `OneCycleLR` can drive it exactly as in the reference -- including the beta1 cycling that
`cycle_momentum=True` applies to Adam-family optimizers: lr and beta1 are read from the
group at every step and passed to the kernels as per-step scalars.

Two ways to consume gradients:
  * `step()`                -- after a normal autograd backward.  A first layer whose gradient was left
                               factored by model.linear._MlpFunction (`weight._vs_lowrank_grad`) is
                               updated by vs_dw_adamw_fused (dW never materialised); every other
                               parameter by vs_adamw.
  * `begin_fused_step()`    -- used by Linear.fused_train_step: bumps the step counters and returns
                               the hyper-parameters; the update itself happens inside vs_mlp_train_step.
"""
from __future__ import annotations

import torch

import vsb200 as vs


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        for group in self.param_groups:
            for p in group["params"]:
                p._vs_defer_ok = True   # lets the first layer keep its gradient factored

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            if not p.is_cuda:
                raise vs.VsError("FusedAdamW only updates CUDA parameters (no CPU fallback)")
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @staticmethod
    def _hyper(group, step):
        b1, b2 = group["betas"]
        return vs.AdamWHyper(float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                             float(group["weight_decay"]), int(step))

    def begin_fused_step(self):
        if len(self.param_groups) != 1:
            raise vs.VsError("fused train step supports a single parameter group (as src/train.py builds)")
        group = self.param_groups[0]
        step = None
        for p in group["params"]:
            st = self._state_for(p)
            st["step"] += 1
            step = st["step"]
        return self._hyper(group, step)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stream = vs.stream()
        for group in self.param_groups:
            for p in group["params"]:
                low = getattr(p, "_vs_lowrank_grad", None)
                if p.grad is None and low is None:
                    continue
                st = self._state_for(p)
                st["step"] += 1
                h = self._hyper(group, st["step"])
                if low is not None:
                    dy, x = low
                    x_f32 = x if x.dtype == torch.float32 else None
                    x_u8 = x if x.dtype == torch.uint8 else None
                    vs.check(vs.lib.vs_dw_adamw_fused(vs.ptr(dy), vs.ptr(x_f32), vs.ptr(x_u8), vs.ptr(p.data),
                                                      vs.ptr(st["exp_avg"]), vs.ptr(st["exp_avg_sq"]), dy.shape[0],
                                                      p.shape[1], p.shape[0], h, stream))
                    p._vs_lowrank_grad = None
                else:
                    g = p.grad.contiguous()
                    vs.check(vs.lib.vs_adamw(vs.ptr(p.data), vs.ptr(g), vs.ptr(st["exp_avg"]), vs.ptr(st["exp_avg_sq"]),
                                             p.numel(), h, stream))
        return loss

    def zero_grad(self, set_to_none: bool = True):
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_vs_lowrank_grad", None) is not None:
                    p._vs_lowrank_grad = None
        super().zero_grad(set_to_none=set_to_none)
