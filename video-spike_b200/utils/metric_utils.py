"""Evaluation metrics with the definitions of the reference's src/utils/metric_utils.py
(neg_log_likelihood :36-76, bits_per_spike :78-102, r2_score :8-15).  They run once per epoch on
the host in float64, exactly like the reference; the hot step never calls them.
"""
from __future__ import annotations

import logging

import numpy as np
import torch
from scipy.special import gammaln

logger = logging.getLogger(__name__)


def r2_score(y_true, y_pred, device="cpu"):
    """torcheval R2Score semantics: 1 - SS_res / SS_tot over all samples (single output)."""
    y_true = torch.as_tensor(y_true, dtype=torch.float64).flatten()
    y_pred = torch.as_tensor(y_pred, dtype=torch.float64).flatten()
    ss_res = torch.sum((y_true - y_pred) ** 2)
    ss_tot = torch.sum((y_true - y_true.mean()) ** 2)
    return float(1.0 - ss_res / ss_tot)


def neg_log_likelihood(rates, spikes, zero_warning=True):
    """Poisson NLL summed over all bins: r - n log r + log n!."""
    assert spikes.shape == rates.shape, \
        f"neg_log_likelihood: Rates and spikes should be of the same shape. spikes: {spikes.shape}, rates: {rates.shape}"
    if np.any(np.isnan(spikes)):
        keep = ~np.isnan(spikes)
        rates, spikes = rates[keep], spikes[keep]
    assert not np.any(np.isnan(rates)), "neg_log_likelihood: NaN rate predictions found"
    assert np.all(rates >= 0), "neg_log_likelihood: Negative rate predictions found"
    if np.any(rates == 0):
        if zero_warning:
            logger.warning("neg_log_likelihood: Zero rate predictions found. Replacing zeros with 1e-9")
        rates[rates == 0] = 1e-9
    return np.sum(rates - spikes * np.log(rates) + gammaln(spikes + 1.0))


def bits_per_spike(rates, spikes):
    """(NLL of the per-neuron mean-rate model - NLL of the prediction) / total spikes / ln 2."""
    nll_model = neg_log_likelihood(rates, spikes)
    lead_axes = tuple(range(spikes.ndim - 1))
    null_rates = np.tile(np.nanmean(spikes, axis=lead_axes, keepdims=True), spikes.shape[:-1] + (1,))
    nll_null = neg_log_likelihood(null_rates, spikes, zero_warning=False)
    return (nll_null - nll_model) / np.nansum(spikes) / np.log(2)


# ---- device versions: same definitions, no copy of the predictions to the host -----------------
def device_bits_per_spike(rates_KTN, spikes_KTN):
    """Per-neuron bits per spike, float64 (N,) CUDA tensor: bits_per_spike(rates[:, :, [n]], spikes[:, :, [n]]) of
    metric_utils.py:78-102 for every n at once (vs_bits_per_spike)."""
    import vsb200 as vs
    r = rates_KTN.contiguous().float()
    s = spikes_KTN.contiguous().float()
    K, T, N = r.shape
    out = torch.empty(N, dtype=torch.float64, device=r.device)
    vs.check(vs.lib.vs_bits_per_spike(vs.ptr(r), vs.ptr(s), K, T, N, vs.ptr(out), vs.stream()))
    return out


def device_r2_per_trial(gt_KTN, pred_KTN):
    """sklearn r2_score(gt[k].T-style (N, T) slices) per trial k, float64 (K,) CUDA tensor (utils.py:158)."""
    import vsb200 as vs
    g = gt_KTN.contiguous().float()
    p = pred_KTN.contiguous().float()
    K, T, N = g.shape
    out = torch.empty(K * T, dtype=torch.float64, device=g.device)
    vs.check(vs.lib.vs_r2_rows(vs.ptr(g), vs.ptr(p), K, T, N, vs.ptr(out), vs.stream()))
    return out.view(K, T).mean(dim=1)
