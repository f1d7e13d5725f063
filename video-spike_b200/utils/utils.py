"""Host-side helpers with the behaviour of the reference's src/utils/utils.py for the symbols the
hot path touches: model registry (:28-34), CLI flags (:36-47), seeding (:49-59), batch device move
(:61-66), z-score statistics (:107-112), one-hot (:114-119) and the per-epoch metric loop (:122-181).
CEBRA / PCA / plotting helpers of that file are outside the hot path and not provided.
"""
from __future__ import annotations

import argparse
import os
import random

import numpy as np
import torch
from sklearn.metrics import r2_score as r2_score_sklearn

from model.linear import Linear
from model.rrr import train_model, train_model_main  # noqa: F401  (re-exported like the reference)
from utils.metric_utils import bits_per_spike

NAME2MODEL = {"Linear": Linear}


def get_args(argv=None):
    p = argparse.ArgumentParser(description="IBL Spike Video Project")
    p.add_argument("--model_config", type=str, default="configs/model/model_config.yaml", help="Model config file")
    p.add_argument("--train_config", type=str, default="configs/train/train_config.yaml", help="Train config file")
    p.add_argument("--seed", type=int, default=42, help="Random seed")
    p.add_argument("--log_dir", type=str, default="logs", help="Log directory")
    p.add_argument("--eid", type=str, default="d57df551-6dcb-4242-9c72-b806cff5613a")
    p.add_argument("--input_mod", type=str, default="whisker-motion-energy", help="Input modality")
    p.add_argument("--model", type=str, default="cm", help="Model name")
    p.add_argument("--save_plot", action="store_true", help="Save plot")
    return p.parse_args(argv)


def set_seed(seed):
    os.environ["PYTHONHASHSEED"] = str(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    print("seed set to {}".format(seed))


def move_batch_to_device(batch, device):
    for key, val in batch.items():
        if isinstance(val, torch.Tensor):
            batch[key] = val.to(device, non_blocking=True)
    return batch


def _std(arr):
    mean = np.mean(arr, axis=0)
    std = np.clip(np.std(arr, axis=0), 1e-8, None)
    return (arr - mean) / std, mean, std


def _one_hot(arr, T):
    levels = np.sort(np.unique(arr))
    out = np.zeros((len(arr), T, len(levels)))
    for i, lvl in enumerate(levels):
        out[:, :, i] = arr == lvl
    return out


def metrics_list(gt, pred, metrics=("bps", "rsquared"), device="cpu"):
    """gt / pred arrive as (N, T, K) (the trainer transposes, src/trainer/base.py:190-191).

    Kept on purpose: both loops run over gt.shape[-1] -- the TRIAL count -- while "bps"
    indexes the neuron axis of the re-transposed arrays, so it covers the first K neurons and raises
    IndexError when K > N; "rsquared" is sklearn's R2 of the (N, T) slice of trial i."""
    results = {}
    if gt.is_cuda and pred.is_cuda and set(metrics) <= {"bps", "rsquared"}:
        # device route (vs_bits_per_spike / vs_r2_rows): same numbers, no per-neuron Python loop, no copy of the predictions
        from utils.metric_utils import device_bits_per_spike, device_r2_per_trial
        g_ktn, p_ktn = gt.transpose(-1, 0).contiguous(), pred.transpose(-1, 0).contiguous()
        K, T, N = g_ktn.shape
        if "bps" in metrics:
            if K > N:
                raise IndexError(f"index {N} is out of bounds for axis 2 with size {N}")        # the reference's quirk (A8)
            bps = device_bits_per_spike(p_ktn, g_ktn)[:K]                                      # first K NEURONS (A8)
            bps = torch.where(torch.isinf(bps), torch.full_like(bps, float("nan")), bps)
            results["bps"] = float(torch.nanmean(bps)) if bool((~torch.isnan(bps)).any()) else float("nan")
        if "rsquared" in metrics:
            r2 = device_r2_per_trial(g_ktn, p_ktn)
            results["rsquared"] = float(torch.nanmean(r2))
        return results
    if "bps" in metrics:
        g = gt.transpose(-1, 0).cpu().numpy()
        p = pred.transpose(-1, 0).cpu().numpy()
        vals = []
        for i in range(gt.shape[-1]):
            bps = bits_per_spike(p[:, :, [i]], g[:, :, [i]])
            vals.append(np.nan if np.isinf(bps) else bps)
        results["bps"] = np.nanmean(vals)
    if "rsquared" in metrics:
        g, p = gt.cpu().clone(), pred.cpu().numpy()
        vals = [r2_score_sklearn(y_true=g[:, :, i], y_pred=p[:, :, i]) for i in range(gt.shape[-1])]
        results["rsquared"] = np.nanmean(vals)
    if "mse" in metrics:
        results["mse"] = torch.mean((gt - pred) ** 2)
    if "mae" in metrics:
        results["mae"] = torch.mean(torch.abs(gt - pred))
    return results


def select_frames(n_total=119, n_keep=100):
    """src/train_rrr.py:48-49: 100 of frames 0..118, drawn from the GLOBAL numpy stream (right after
    `set_seed(config.seed)` in the reference driver), returned sorted."""
    idx = np.random.choice(n_total, n_keep, replace=False)
    return np.sort(idx)


def evaluate_rrr_session(pred, gt_held_out, threshold=1e-3):
    """src/train_rrr.py:198-224 and src/utils/utils.py:417-447: clip the de-z-scored prediction at 1e-3, per-neuron
    bits/spike against the held-out counts, mean over trials of sklearn's R2 per neuron (inf bps -> nan)."""
    from tqdm import tqdm
    pred = np.clip(pred, threshold, None)
    bps_list, r2_list = [], []
    for n_i in tqdm(range(pred.shape[2]), desc='co-bps'):
        bps = bits_per_spike(pred[:, :, [n_i]], gt_held_out[:, :, [n_i]])
        r2 = np.nanmean([r2_score_sklearn(gt_held_out[k, :, n_i], pred[k, :, n_i]) for k in range(pred.shape[0])])
        r2_list.append(r2)
        bps_list.append(np.nan if np.isinf(bps) else bps)
    return pred, bps_list, r2_list


def _zscore_session_inplace(entry):
    """z-score X and y of one session with the TRAIN statistics (std clipped at 1e-8), append the ones column and record
    the statistics under entry["setup"] -- the per-session body of src/utils/utils.py:378-396.  Returns the raw test y."""
    _, mean_X, std_X = _std(entry["X"][0])
    _, mean_y, std_y = _std(entry["y"][0])
    held_out = entry["y"][1].copy()
    for split in range(2):
        X = (entry["X"][split] - mean_X) / std_X
        K, T = entry["X"][split].shape[0], entry["X"][split].shape[1]
        if X.ndim == 2:                                   # the reference expands a 2-D array to (1, K, T) here; kept
            X = np.expand_dims(X, axis=0)
        entry["X"][split] = np.concatenate([X, np.ones((K, T, 1))], axis=2)
        entry["y"][split] = (entry["y"][split] - mean_y) / std_y
        print(f"X shape: {entry['X'][split].shape}, y shape: {entry['y'][split].shape}")
    entry["setup"].update(mean_X_Tv=mean_X, std_X_Tv=std_X, mean_y_TN=mean_y, std_y_TN=std_y)
    return held_out


def train_rrr(data_dict):
    """Drop-in for src/utils/utils.py:376-456 -- the RRR fit that ContrastTrainer._validate runs every validation round
    (src/trainer/contrast.py:129-162): per session, z-score with the train statistics, ones column, an independent fit
    with l2 = 100 and rank 3, and the co-bps / R2 of the de-z-scored test prediction.  `data_dict` is mutated in place
    like in the reference and the result dictionary has the same keys."""
    held_out = {eid: _zscore_session_inplace(data_dict[eid]) for eid in data_dict}
    print("Training RRR")
    result = {}
    for eid in data_dict:
        model, _ = train_model_main(train_data={eid: data_dict[eid]}, l2=100, n_comp=3, model_fname='tmp', save=False)
        print(f"Model {eid} trained")
        with torch.no_grad():
            pred_fr = model.predict_y_fr(data_dict, eid, 1)[2]
        pred, bps_list, r2_list = evaluate_rrr_session(pred_fr.cpu().numpy(), held_out[eid])
        print(f"Co-BPS: {np.nanmean(bps_list)}")
        print(f"r2: {np.nanmean(r2_list)}")
        result[eid] = {'gt': held_out[eid], 'pred': pred, 'bps': bps_list, 'r2': r2_list, 'eid': eid}
    return result
