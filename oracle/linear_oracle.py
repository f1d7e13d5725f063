"""CPU oracle for the `Linear` (MLP) train path.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Restates, with explicit formulas on torch-CPU float32 tensors (no autograd, no nn.Module):
  src/loader/base.py:39,54                frames.float() -- no scaling
  src/trainer/base.py:61-70               flatten(1) + concat of the input modalities
  src/model/linear.py:10-15,24-32,45-53   6 Linear layers, ReLU after layers 0,1,3,4, reshape (B,100,N)
  src/train.py:59, trainer/base.py:141-143  PoissonNLLLoss(log_input=True, reduction='none').mean()
  src/trainer/base.py:150                 backward (written out by hand below)
  src/train.py:44-49                      torch.optim.AdamW (third-party; restated from its documented update rule)
  src/train.py:51-57                      OneCycleLR, cos annealing, three_phase=False, cycle_momentum=True
Pinned by tests/golden/linear_*.npz generated from the reference's own `model.linear.Linear`
driven by torch.optim.AdamW / OneCycleLR (oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

RELU_AFTER = (True, True, False, True, True, False)   # linear.py:24-32 (encoder), 45-53 (decoder)
STATE_KEYS = ("encoder.layers.0", "encoder.layers.2", "encoder.layers.4",
              "decoder.layers.0", "decoder.layers.2", "decoder.layers.4")


def init_params(input_dim: int, n_neurons: int, seed: int = 42, enc_hidden=(256, 128), bottleneck=64,
                dec_hidden=(128, 256)):
    """Default nn.Linear initialisation in the reference's construction order (encoder then decoder,
    linear.py:6-7) under torch.manual_seed(seed) (seed=None: continue the current global stream, which is what
    src/train.py does -- the DataLoader iterator created by get_metadata_from_loader, train.py:35, draws its base seed
    from torch's global generator BEFORE the model is built).  Returns [(W, b)] in forward order."""
    if seed is not None:
        torch.manual_seed(seed)
    dims = [input_dim, *enc_hidden, bottleneck, *dec_hidden, 100 * n_neurons]
    params = []
    for i in range(len(dims) - 1):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        params.append((lin.weight.detach().clone(), lin.bias.detach().clone()))
    return params


def to_state_dict(params):
    sd = {}
    for key, (W, b) in zip(STATE_KEYS, params):
        sd[key + ".weight"], sd[key + ".bias"] = W, b
    return sd


def cast_frames(frames_u8: torch.Tensor) -> torch.Tensor:
    """loader/base.py:39 + trainer/base.py:66: uint8 -> float32 (0..255 kept), flatten(1)."""
    return frames_u8.float().flatten(1)


def forward(params, x: torch.Tensor, keep=False):
    acts = []
    h = x
    for (W, b), relu in zip(params, RELU_AFTER):
        h = h @ W.t() + b
        if relu:
            h = torch.clamp_min(h, 0.0)
        acts.append(h)
    return (h, acts) if keep else h


def poisson_nll_mean(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return (torch.exp(logits) - target * logits).mean()


def loss_and_grads(params, x: torch.Tensor, target_flat: torch.Tensor):
    """Forward + hand-written backward.  Returns (loss, [(dW, db)])."""
    logits, acts = forward(params, x, keep=True)
    n = logits.numel()
    loss = poisson_nll_mean(logits, target_flat)
    g = (torch.exp(logits) - target_flat) / n
    grads = [None] * len(params)
    for l in range(len(params) - 1, -1, -1):
        if RELU_AFTER[l]:
            g = g * (acts[l] > 0).to(g.dtype)
        inp = x if l == 0 else acts[l - 1]
        grads[l] = (g.t() @ inp, g.sum(0))
        if l > 0:
            g = g @ params[l][0]
    return loss, grads


def adamw_update(p, g, m, v, step: int, lr: float, beta1: float, beta2: float = 0.999, eps: float = 1e-8,
                 wd: float = 0.01):
    """torch.optim.AdamW (decoupled weight decay, bias correction), in place on float32 tensors."""
    p.mul_(1.0 - lr * wd)
    m.lerp_(g, 1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def one_cycle(step: int, total_steps: int, max_lr: float, pct_start: float, div_factor: float,
              final_div_factor: float = 1e4, base_momentum: float = 0.85, max_momentum: float = 0.95):
    """lr and beta1 that OneCycleLR leaves in the optimizer for optimizer step number `step`
    (0-based: step 0 is the value set at scheduler construction)."""
    initial_lr = max_lr / div_factor
    min_lr = initial_lr / final_div_factor
    end1 = float(pct_start * total_steps) - 1.0
    end2 = float(total_steps) - 1.0

    def cos(start, end, pct):
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    if step <= end1:
        pct = step / end1
        return cos(initial_lr, max_lr, pct), cos(max_momentum, base_momentum, pct)
    pct = (step - end1) / (end2 - end1)
    return cos(max_lr, min_lr, pct), cos(base_momentum, max_momentum, pct)


class Trainer:
    """The step body of src/trainer/base.py:147-154 on the oracle's tensors."""

    def __init__(self, params, total_steps=5000, max_lr=5e-5, wd=0.01, eps=1e-8, pct_start=0.15, div_factor=10.0):
        self.params = [(W.clone(), b.clone()) for W, b in params]
        self.m = [(torch.zeros_like(W), torch.zeros_like(b)) for W, b in params]
        self.v = [(torch.zeros_like(W), torch.zeros_like(b)) for W, b in params]
        self.t = 0
        self.cfg = dict(total_steps=total_steps, max_lr=max_lr, pct_start=pct_start, div_factor=div_factor)
        self.wd, self.eps = wd, eps

    def hyper(self):
        return one_cycle(self.t, **self.cfg)

    def step(self, frames_u8: torch.Tensor, target: torch.Tensor) -> float:
        x = cast_frames(frames_u8)
        loss, grads = loss_and_grads(self.params, x, target.reshape(target.shape[0], -1).float())
        lr, beta1 = self.hyper()
        self.t += 1
        for (W, b), (gW, gb), (mW, mb), (vW, vb) in zip(self.params, grads, self.m, self.v):
            adamw_update(W, gW, mW, vW, self.t, lr, beta1, eps=self.eps, wd=self.wd)
            adamw_update(b, gb, mb, vb, self.t, lr, beta1, eps=self.eps, wd=self.wd)
        return float(loss)

    def predict_rates(self, frames_u8: torch.Tensor, n_neurons: int) -> torch.Tensor:
        """eval path, src/trainer/base.py:172-186: exp of the logits, (B,100,N)."""
        return torch.exp(forward(self.params, cast_frames(frames_u8))).reshape(-1, 100, n_neurons)


def synth_batch(batch: int, frame_shape, n_neurons: int, seed: int = 0, dist: str = "sparse"):
    """Synthetic inputs of the loader's shape.  dist='sparse' is the distribution the
    reference is stable on (randint(1,9) * Bernoulli(0.01)); 'uniform' is 0..255 for throughput."""
    g = torch.Generator().manual_seed(seed)
    shape = (batch, *frame_shape)
    if dist == "uniform":
        frames = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    else:
        frames = (torch.randint(1, 9, shape, generator=g, dtype=torch.uint8)
                  * (torch.rand(shape, generator=g) < 0.01).to(torch.uint8))
    ap = torch.poisson(torch.full((batch, 100, n_neurons), 0.3), generator=g)
    return frames, ap
