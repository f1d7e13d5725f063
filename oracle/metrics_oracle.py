"""CPU oracle for the evaluation metrics.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Restates
  src/utils/metric_utils.py:36-76    neg_log_likelihood
  src/utils/metric_utils.py:78-102   bits_per_spike
  src/utils/utils.py:122-181         metrics_list ("bps" and "rsquared" branches, with
                                     the trial-count-indexes-neurons quirk)
  sklearn.metrics.r2_score           (third-party; restated: 1 - SS_res/SS_tot per output,
                                     uniform average, force_finite semantics)
Pinned by tests/golden/metrics_kat.npz (values produced by the reference's own
metric_utils imported with a torcheval stub, oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln


def neg_log_likelihood(rates: np.ndarray, spikes: np.ndarray) -> float:
    """src/utils/metric_utils.py:36-76.  Poisson NLL summed: r - n log r + lgamma(n+1).
    NaN spikes are masked; zero rates become 1e-9 (the reference then logs through an
    undefined `logger`, SURVEY A9 -- the oracle raises NameError the same way only when
    asked to, see `strict_zero`)."""
    assert spikes.shape == rates.shape
    rates = np.array(rates, dtype=np.float64, copy=True)
    spikes = np.asarray(spikes, dtype=np.float64)
    if np.any(np.isnan(spikes)):
        mask = np.isnan(spikes)
        rates, spikes = rates[~mask], spikes[~mask]
    assert not np.any(np.isnan(rates)), "NaN rate predictions found"
    assert np.all(rates >= 0), "Negative rate predictions found"
    rates[rates == 0] = 1e-9
    return float(np.sum(rates - spikes * np.log(rates) + gammaln(spikes + 1.0)))


def bits_per_spike(rates: np.ndarray, spikes: np.ndarray) -> float:
    """src/utils/metric_utils.py:78-102."""
    rates = np.asarray(rates, dtype=np.float64)
    spikes = np.asarray(spikes, dtype=np.float64)
    nll_model = neg_log_likelihood(rates, spikes)
    null_rates = np.tile(np.nanmean(spikes, axis=tuple(range(spikes.ndim - 1)), keepdims=True),
                         spikes.shape[:-1] + (1,))
    nll_null = neg_log_likelihood(null_rates, spikes)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float((nll_null - nll_model) / np.nansum(spikes) / np.log(2))


def r2_score_1d(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """sklearn.metrics.r2_score for one output (force_finite=True)."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    num = np.sum((y_true - y_pred) ** 2)
    den = np.sum((y_true - y_true.mean()) ** 2)
    if den == 0.0:
        return 1.0 if num == 0.0 else 0.0
    return float(1.0 - num / den)


def r2_score_multi(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """sklearn.metrics.r2_score for (n_samples, n_outputs), multioutput='uniform_average'."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    return float(np.mean([r2_score_1d(y_true[:, j], y_pred[:, j]) for j in range(y_true.shape[1])]))


def metrics_list(gt_KTN: np.ndarray, pred_KTN: np.ndarray) -> dict:
    """src/trainer/base.py:188-195 + src/utils/utils.py:125-134,153-167.

    The trainer hands metrics_list gt.transpose(-1,0) = (N,T,K).  Both loops then run
    over range(K) -- the TRIAL count -- while "bps" indexes the neuron axis, so bps covers
    the first K neurons (IndexError if K > N) and "rsquared" scores, per trial, an (N,T)
    slice with time bins as the outputs."""
    K, T, N = gt_KTN.shape
    bps_list = []
    for i in range(K):
        if i >= N:
            raise IndexError("index %d is out of bounds for axis 2 with size %d" % (i, N))
        bps = bits_per_spike(pred_KTN[:, :, [i]], gt_KTN[:, :, [i]])
        bps_list.append(np.nan if np.isinf(bps) else bps)
    r2_list = [r2_score_multi(gt_KTN[i].T, pred_KTN[i].T) for i in range(K)]
    return {"bps": float(np.nanmean(bps_list)), "rsquared": float(np.nanmean(r2_list))}
