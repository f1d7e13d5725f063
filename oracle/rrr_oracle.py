"""CPU oracle for the reduced-rank-regression (RRR) path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py: never imported by the product path).

Restates, in numpy/torch-CPU float64, what the reference does in
  src/train_rrr.py:41,48-49     frame selection (100 of frames 0..118)
  src/train_rrr.py:108-171      smoothing, z-scoring, ones column, X[:, sorted_idx]
  src/utils/utils.py:107-119    _std, _one_hot
  src/model/rrr.py:30-54        RRRGD parameter init
  src/model/rrr.py:79-155       beta, predict, MSE, L2 penalty
  src/model/rrr.py:164-202      one LBFGS.step over the full train split
  src/train_rrr.py:193-236      de-z-scored prediction, clip, co-bps and per-trial R2
torch.optim.LBFGS (third-party, torch==2.2.1 pinned in the reference's env.yaml)
is restated in `lbfgs_step` from its published algorithm (no line search).

Parity is pinned by tests/golden/rrr_*.npz, produced by oracle/make_golden.py from
the reference's own `model.rrr` module imported unmodified.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d

from .metrics_oracle import bits_per_spike, r2_score_1d


# --------------------------------------------------------------------------- R0
def set_seed(seed: int) -> None:
    """src/utils/utils.py:49-59 (the parts that influence CPU results)."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def select_frames(seed: int = 42) -> np.ndarray:
    """src/train_rrr.py:41,48-49: first numpy draw after set_seed(seed)."""
    set_seed(seed)
    idx = np.random.choice(119, 100, replace=False)
    return np.sort(idx)


def zscore_stats(arr: np.ndarray):
    """src/utils/utils.py:107-112 (_std): mean/std over trials, std clipped at 1e-8."""
    mean = np.mean(arr, axis=0)
    std = np.clip(np.std(arr, axis=0), 1e-8, None)
    return mean, std


def one_hot(arr: np.ndarray, T: int) -> np.ndarray:
    """src/utils/utils.py:114-119 (_one_hot)."""
    uni = np.sort(np.unique(arr))
    ret = np.zeros((len(arr), T, len(uni)))
    for i, u in enumerate(uni):
        ret[:, :, i] = arr == u
    return ret


def preprocess_session(X_splits, y_splits, sorted_idx, smooth_w: float = 2.0):
    """src/train_rrr.py:108-171 for the video-like modalities (no one-hot branch).

    X_splits: [train, test] arrays (K, 120, F) (any real dtype); y_splits: [train, test]
    arrays (K, 100, N) spike counts.  Returns (data, ground_truth) where data has the
    layout RRRGD expects: {"X": [..], "y": [..], "setup": {...}} and ground_truth is the
    unsmoothed test y (train_rrr.py:116).
    """
    ground_truth = y_splits[1]
    y_s = [gaussian_filter1d(np.asarray(y, dtype=np.float64), smooth_w, axis=1) for y in y_splits]
    X_s = [np.asarray(x, dtype=np.float64) for x in X_splits]
    mean_X, std_X = zscore_stats(X_s[0])
    mean_y, std_y = zscore_stats(y_s[0])
    Xo, yo = [], []
    for i in range(2):
        K, T = X_s[i].shape[0], X_s[i].shape[1]
        x = (X_s[i] - mean_X) / std_X
        x = np.concatenate([x, np.ones((K, T, 1))], axis=2)
        x = x[:, sorted_idx]
        Xo.append(x)
        yo.append((y_s[i] - mean_y) / std_y)
    data = {"X": Xo, "y": yo,
            "setup": {"mean_X_Tv": mean_X, "std_X_Tv": std_X, "mean_y_TN": mean_y, "std_y_TN": std_y}}
    return data, ground_truth


# --------------------------------------------------------------------------- R1
def rrr_init(train_data: dict, ncomp: int):
    """src/model/rrr.py:35-49.  Returns ({name: float64 array}, order) with the
    ParameterDict insertion order ({eid}_U, {eid}_b ..., V)."""
    np.random.seed(0)
    params = {}
    V = None
    for eid in train_data:
        _X = train_data[eid]["X"][0]
        _y = train_data[eid]["y"][0]
        K, T, ncoef = _X.shape
        _, _, N = _y.shape
        U = np.random.normal(size=(N, ncoef - 1, ncomp)) / np.sqrt(T * ncomp)
        V = np.random.normal(size=(ncomp, T)) / np.sqrt(T * ncomp)
        b = np.ascontiguousarray(np.expand_dims(_y.mean(0).T, 1))
        params[f"{eid}_U"] = U
        params[f"{eid}_b"] = b
    params["V"] = V
    return params


# ---------------------------------------------------------------------- R2 - R4
def compute_beta(U, V, b):
    """src/model/rrr.py:79-96: beta = cat(U @ V, b, dim=1) -> (N, ncoef, T)."""
    return np.concatenate([U @ V, b], axis=1)


def _einsum(eq, *ops):
    """torch.einsum on CPU float64 (the same routine the reference calls, rrr.py:113; it maps to
    multi-threaded bmm, which also makes this the fair host-core baseline for bench.py)."""
    return torch.einsum(eq, *[torch.from_numpy(np.ascontiguousarray(o)) for o in ops]).numpy()


def predict(beta, X):
    """src/model/rrr.py:105-116: einsum('ktc,nct->ktn')."""
    return _einsum("ktc,nct->ktn", X, beta)


def loss_and_grad_dense(params: dict, data: dict, l2: float, split: int = 0):
    """Closure of src/model/rrr.py:165-175 with analytic gradients (beta materialised,
    exactly the reference's formulation).  Returns (loss, grads dict, per-eid SSE (N,))."""
    V = params["V"]
    grads = {"V": np.zeros_like(V)}
    loss = 0.0
    sse = {}
    for eid in data:
        U, b = params[f"{eid}_U"], params[f"{eid}_b"]
        X, y = data[eid]["X"][split], data[eid]["y"][split]
        beta = compute_beta(U, V, b)                       # (N, C, T)
        R = predict(beta, X) - y                           # (K, T, N)
        sse[eid] = np.sum(R ** 2, axis=(0, 1))
        loss += sse[eid].sum() + l2 * np.sum(beta ** 2)
        dbeta = 2.0 * _einsum("ktc,ktn->nct", X, R) + 2.0 * l2 * beta
        grads[f"{eid}_U"] = dbeta[:, :-1, :] @ V.T         # (N, C-1, r)
        grads[f"{eid}_b"] = dbeta[:, -1:, :]
        grads["V"] += _einsum("ncj,nct->jt", U, dbeta[:, :-1, :])
    return loss, grads, sse


def loss_and_grad_autograd(params: dict, data: dict, l2: float, split: int = 0):
    """The reference closure op for op (src/model/rrr.py:147-155,165-175) on torch-CPU float64 with
    autograd: beta is built twice (MSE + penalty), X/y are converted on every call, einsum + backward.
    This is the variant bench.py times as the host-core baseline; it equals loss_and_grad_dense."""
    tp = {k: torch.from_numpy(np.ascontiguousarray(v)).clone().requires_grad_(True) for k, v in params.items()}
    total = 0.0
    for eid in data:
        beta = torch.cat((tp[f"{eid}_U"] @ tp["V"], tp[f"{eid}_b"]), 1)
        X = torch.from_numpy(data[eid]["X"][split])
        y = torch.from_numpy(data[eid]["y"][split])
        ypred = torch.einsum("ktc,nct->ktn", X, beta)
        total = total + torch.sum((ypred - y) ** 2, axis=(0, 1)).sum()
        beta2 = torch.cat((tp[f"{eid}_U"] @ tp["V"], tp[f"{eid}_b"]), 1)
        total = total + l2 * torch.sum(beta2 ** 2)
    total.backward()
    return float(total), {k: v.grad.numpy() for k, v in tp.items()}


def _bf16(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float64).numpy()


def loss_and_grad_lowrank(params: dict, data: dict, l2: float, split: int = 0, emulate_bf16: bool = False):
    """Same closure in the factorised form the CUDA kernels use (DESIGN.md "RRR closure"):
        Z[k,t,n,j] = sum_c X[k,t,c] U[n,c,j];   yhat = sum_j V[j,t] Z[..,j] + x_last * b
        dU = 2 X^T (R (x) V) + 2 l2 U (V V^T);  dV = 2 sum_{k,n} R Z + 2 l2 (U^T U) V
    With emulate_bf16 the three tensor-core operands (X, U, R(x)V) are rounded to bf16 the
    way the kernels round them; accumulation stays wide."""
    V = params["V"]
    r, T = V.shape
    W = V @ V.T
    grads = {"V": np.zeros_like(V)}
    loss = 0.0
    sse = {}
    for eid in data:
        U, b = params[f"{eid}_U"], params[f"{eid}_b"]
        X, y = data[eid]["X"][split], data[eid]["y"][split]
        K, _, C = X.shape
        N = U.shape[0]
        Xm, xl = X[:, :, :-1], X[:, :, -1]                 # (K,T,C'), (K,T)
        Xq = _bf16(Xm) if emulate_bf16 else Xm
        Uq = _bf16(U) if emulate_bf16 else U
        Z = (Xq.reshape(K * T, C - 1) @ Uq.transpose(1, 0, 2).reshape(C - 1, N * r)).reshape(K, T, N, r)
        yhat = np.einsum("ktnj,jt->ktn", Z, V) + xl[:, :, None] * b[:, 0, :].T[None]
        R = yhat - y
        sse[eid] = np.sum(R ** 2, axis=(0, 1))
        G = np.einsum("nci,ncj->ij", U, U)
        loss += sse[eid].sum() + l2 * (np.sum(G * W) + np.sum(b ** 2))
        RV = np.einsum("ktn,jt->ktnj", R, V)
        RVq = _bf16(RV) if emulate_bf16 else RV
        dU = 2.0 * (Xq.reshape(K * T, C - 1).T @ RVq.reshape(K * T, N * r)).reshape(C - 1, N, r).transpose(1, 0, 2)
        grads[f"{eid}_U"] = dU + 2.0 * l2 * (U @ W)
        grads[f"{eid}_b"] = (2.0 * np.einsum("kt,ktn->nt", xl, R) + 2.0 * l2 * b[:, 0, :])[:, None, :]
        grads["V"] += 2.0 * np.einsum("ktn,ktnj->jt", R, Z) + 2.0 * l2 * (G @ V)
    return loss, grads, sse


# --------------------------------------------------------------------------- R5
def lbfgs_step(closure, x0: np.ndarray, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7,
               tolerance_change=1e-9, history_size=100):
    """One torch.optim.LBFGS.step (line_search_fn=None), restated from its published
    algorithm (torch/optim/lbfgs.py, `step`).  closure(x) -> (loss, flat_grad).
    Returns (x, trace) where trace lists (loss, |g|_inf) per closure evaluation."""
    if max_eval is None:
        max_eval = max_iter * 5 // 4
    x = x0.copy()
    loss, g = closure(x)
    trace = [(float(loss), float(np.abs(g).max()))]
    evals = 1
    if np.abs(g).max() <= tolerance_grad:
        return x, trace
    old_dirs, old_stps, ro = [], [], []
    H_diag = 1.0
    d = t = prev_g = None
    n_iter = 0
    while n_iter < max_iter:
        n_iter += 1
        if n_iter == 1:
            d = -g
        else:
            yv = g - prev_g
            s = d * t
            ys = float(yv @ s)
            if ys > 1e-10:
                if len(old_dirs) == history_size:
                    old_dirs.pop(0); old_stps.pop(0); ro.pop(0)
                old_dirs.append(yv); old_stps.append(s); ro.append(1.0 / ys)
                H_diag = ys / float(yv @ yv)
            num_old = len(old_dirs)
            al = [0.0] * num_old
            q = -g
            for i in range(num_old - 1, -1, -1):
                al[i] = float(old_stps[i] @ q) * ro[i]
                q = q - al[i] * old_dirs[i]
            d = q * H_diag
            for i in range(num_old):
                be_i = float(old_dirs[i] @ d) * ro[i]
                d = d + (al[i] - be_i) * old_stps[i]
        prev_g = g.copy()
        prev_loss = loss
        t = min(1.0, 1.0 / float(np.abs(g).sum())) * lr if n_iter == 1 else lr
        gtd = float(g @ d)
        if gtd > -tolerance_change:
            break
        x = x + t * d
        ls_evals = 0
        if n_iter != max_iter:
            loss, g = closure(x)
            trace.append((float(loss), float(np.abs(g).max())))
            ls_evals = 1
        evals += ls_evals
        if n_iter == max_iter:
            break
        if evals >= max_eval:
            break
        if np.abs(g).max() <= tolerance_grad:
            break
        if np.abs(d * t).max() <= tolerance_change:
            break
        if abs(loss - prev_loss) < tolerance_change:
            break
    return x, trace


def _flatten(params: dict, order):
    return np.concatenate([params[k].ravel() for k in order])


def _unflatten(x: np.ndarray, like: dict, order):
    out, o = {}, 0
    for k in order:
        n = like[k].size
        out[k] = x[o:o + n].reshape(like[k].shape)
        o += n
    return out


def train_model_main(train_data: dict, l2: float, n_comp: int, lowrank: bool = False, emulate_bf16: bool = False):
    """src/model/rrr.py:164-202: init, one LBFGS.step on split 0, val SSE on split 1.
    Returns (params, {"mses_val": {eid: (N,)}, "mse_val_mean": float}, trace)."""
    params = rrr_init(train_data, n_comp)
    order = list(params.keys())
    fn = (lambda p, s: loss_and_grad_lowrank(p, train_data, l2, s, emulate_bf16)) if lowrank else \
         (lambda p, s: loss_and_grad_dense(p, train_data, l2, s))

    def closure(x):
        p = _unflatten(x, params, order)
        loss, grads, _ = fn(p, 0)
        return loss, _flatten(grads, order)

    x, trace = lbfgs_step(closure, _flatten(params, order))
    params = _unflatten(x, params, order)
    _, _, sse_val = fn(params, 1)
    return params, {"mses_val": sse_val, "mse_val_mean": float(sum(v.sum() for v in sse_val.values()))}, trace


# --------------------------------------------------------------------------- R6
def predict_y_fr(params: dict, data: dict, eid: str, split: int):
    """src/model/rrr.py:122-142: prediction mapped back to firing-rate units."""
    beta = compute_beta(params[f"{eid}_U"], params["V"], params[f"{eid}_b"])
    X, y = data[eid]["X"][split], data[eid]["y"][split]
    yhat = predict(beta, X)
    mean_y, std_y = data[eid]["setup"]["mean_y_TN"], data[eid]["setup"]["std_y_TN"]
    return X, y * std_y + mean_y, yhat * std_y + mean_y


def eval_session(pred: np.ndarray, gt_held_out: np.ndarray, threshold: float = 1e-3):
    """src/train_rrr.py:193-236: clip, per-neuron bits/spike against the unsmoothed
    test counts, per-trial R2 averaged with nanmean."""
    pred = np.clip(pred, threshold, None)
    bps_list, r2_list = [], []
    for n_i in range(pred.shape[2]):
        bps = bits_per_spike(pred[:, :, [n_i]], gt_held_out[:, :, [n_i]])
        r2s = [r2_score_1d(gt_held_out[k, :, n_i], pred[k, :, n_i]) for k in range(pred.shape[0])]
        r2_list.append(np.nanmean(r2s))
        bps_list.append(np.nan if np.isinf(bps) else bps)
    return {"co_bps": float(np.nanmean(bps_list)), "r2": float(np.nanmean(r2_list)),
            "bps_list": bps_list, "r2_list": r2_list, "pred": pred}
