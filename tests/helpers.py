"""Shared builders for the tests, smoke() and bench.py (no oracle use here)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "video-spike_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def linear_config(input_dim, n_neurons):
    """config.model exactly as src/train.py:36-42 leaves it for config/model/linear_video.yaml."""
    from utils.config_utils import config_from_kwargs, update_config
    cfg = config_from_kwargs({"model": "include:" + os.path.join(PKG, "config", "model", "linear_video.yaml")})
    cfg = update_config(os.path.join(PKG, "config", "train", "linear_video.yaml"), cfg)
    cfg["model"]["encoder"]["input_dim"] = input_dim
    cfg["model"]["decoder"]["output_dim"] = 100 * n_neurons
    return cfg


def make_linear_model(input_dim, n_neurons, device, total_steps=5000, seed=42, fused=True):
    """Model + optimizer + OneCycleLR built the way src/train.py:39-57 builds them."""
    from model.linear import Linear
    from optim import FusedAdamW
    cfg = linear_config(input_dim, n_neurons)
    torch.manual_seed(seed)
    model = Linear(cfg.model).to(device)
    o = cfg.optimizer
    cls = FusedAdamW if fused else torch.optim.AdamW
    opt = cls(model.parameters(), lr=o.lr, weight_decay=o.wd, eps=o.eps)
    sched = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, total_steps=total_steps, max_lr=o.lr,
                                                pct_start=o.warmup_pct, div_factor=o.div_factor)
    return model, opt, sched


def small_rrr_problem(seed=0, K=40, Kt=12, F=96, N=10, eid="e1", raw=False):
    """Synthetic session with a planted rank-3 signal, preprocessed as src/train_rrr.py:108-171 does
    (numpy, float64).  Returns the train_data dict RRRGD expects (or the raw pieces with raw=True)."""
    from scipy.ndimage import gaussian_filter1d
    rng = np.random.default_rng(seed)
    Xtr = rng.integers(0, 256, size=(K, 120, F)).astype(np.uint8)
    Xte = rng.integers(0, 256, size=(Kt, 120, F)).astype(np.uint8)
    Wt = rng.standard_normal((F, 3)) / np.sqrt(F)
    Vt = rng.standard_normal((3, 120))
    A = rng.standard_normal((3, N))

    def rates(X):
        z = ((X.astype(float) - 127.5) / 74.0) @ Wt
        lat = np.einsum("ktj,jt->ktj", z, Vt)[:, :100]
        return np.exp(0.6 * np.einsum("ktj,jn->ktn", lat, A) - 1.0)

    ytr = rng.poisson(rates(Xtr)).astype(np.float64)
    yte = rng.poisson(rates(Xte)).astype(np.float64)
    st = np.random.get_state()
    np.random.seed(42)
    sorted_idx = np.sort(np.random.choice(119, 100, replace=False))
    np.random.set_state(st)
    if raw:
        return Xtr, Xte, ytr, yte, sorted_idx
    ys = [gaussian_filter1d(y, 2, axis=1) for y in (ytr, yte)]
    Xs = [Xtr.astype(np.float64), Xte.astype(np.float64)]
    mean_X, std_X = Xs[0].mean(0), np.clip(Xs[0].std(0), 1e-8, None)
    mean_y, std_y = ys[0].mean(0), np.clip(ys[0].std(0), 1e-8, None)
    Xo = [np.concatenate([(x - mean_X) / std_X, np.ones(x.shape[:2] + (1,))], axis=2)[:, sorted_idx] for x in Xs]
    yo = [(y - mean_y) / std_y for y in ys]
    return {eid: {"X": Xo, "y": yo, "setup": {"mean_X_Tv": mean_X, "std_X_Tv": std_X, "mean_y_TN": mean_y,
                                              "std_y_TN": std_y}, "gt": yte}}
