"""RRR driver: drop-in for the reference's src/train_rrr.py (same CLI, same data dict, same printed metrics,
same `<input_mod>_result.npy`), with the arithmetic on the B200.

    python train_rrr.py --model_config <yaml> --train_config <yaml> --input_mod whisker-video [--eid ...]

Data: `data/data_rrr_<input_mod>.npy` (a pickled dict {eid: {"X": [train, test], "y": [train, test], "setup": {}}},
exactly what the reference loads, train_rrr.py:106) -- or, for the video modalities, uint8 frames
(K, 120, 1, H, W) in the same dict, in which case the whole R0 preprocessing runs on the device
(model.rrr.pack_session_from_frames).  `--input_mod synthetic` builds a seeded synthetic session of the
loader's shape (no dataset ships with this repo).

Multi-GPU: sessions are fitted independently (train_rrr.py:179-187), so under torchrun / accelerate launch
rank r takes sessions r, r+WORLD_SIZE, ... and rank 0 gathers the per-session results; no data-path collective.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d
from sklearn.metrics import r2_score
from tqdm import tqdm

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from model.rrr import pack_session_from_frames, train_model_main  # noqa: E402
from utils.config_utils import config_from_kwargs, update_config  # noqa: E402
from utils.metric_utils import bits_per_spike  # noqa: E402
from utils.utils import _one_hot, _std, evaluate_rrr_session, get_args, select_frames, set_seed  # noqa: E402

VIDEO_LIKE = ['cebra', 'pca', 'ws', 'whisker-video', 'vit', 'cm', 'm', 'c', 'synthetic']


def synthetic_sessions(n_sessions=1, K=64, Kt=16, hw=(22, 33), N=24, seed=0):
    """Seeded stand-in for data_rrr_whisker-video: uint8 frames + Poisson counts with a planted rank-3 signal."""
    rng = np.random.default_rng(seed)
    out = {}
    F = hw[0] * hw[1]
    for s in range(n_sessions):
        Wt = rng.standard_normal((F, 3)) / np.sqrt(F)
        Vt = rng.standard_normal((3, 120))
        A = rng.standard_normal((3, N))
        X, y = [], []
        for k in (K, Kt):
            fr = rng.integers(0, 256, size=(k, 120, 1) + hw, dtype=np.uint8)
            z = ((fr.reshape(k, 120, F).astype(np.float64) - 127.5) / 74.0) @ Wt
            lat = np.einsum("ktj,jt->ktj", z, Vt)[:, :100]
            y.append(rng.poisson(np.exp(0.6 * np.einsum("ktj,jn->ktn", lat, A) - 1.0)).astype(np.float64))
            X.append(fr)
        out[f"synthetic{s:02d}"] = {"X": X, "y": y, "setup": {}}
    return out


def load_train_data(input_mod):
    if input_mod == 'synthetic':
        return synthetic_sessions()
    path = f'data/data_rrr_{input_mod}.npy'
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found (the reference reads it at src/train_rrr.py:106); "
                                "use --input_mod synthetic for a seeded stand-in")
    return np.load(path, allow_pickle=True).item()


def preprocess_host(train_data, eids, input_mod, sorted_idx, smooth_w=2):
    """src/train_rrr.py:108-171 for float inputs (numpy, float64) -- kept on the host because these modalities are
    a few columns wide; the fit itself still runs on the device."""
    for eid in eids:
        for i in range(2):
            train_data[eid]["y"][i] = gaussian_filter1d(train_data[eid]["y"][i], smooth_w, axis=1)
            if input_mod in VIDEO_LIKE:
                if input_mod == 'm':
                    train_data[eid]["X"][i] = train_data[eid]["X"][i][..., :3]
                continue
            if input_mod != 'me' and input_mod != 'of-2d':
                inp = train_data[eid]["X"][i]
                choice, block = inp[:, 0, -2:-1], inp[:, 0, -1:]
                const = 3 if input_mod in ('me-all', 'of-all') else 2
                contin_dim = inp.shape[2] - const
                train_data[eid]["X"][i] = np.concatenate([_one_hot(choice, 120), _one_hot(block, 120), inp[..., -2 - contin_dim:-2]], axis=2)
    for eid in eids:
        _, mean_X, std_X = _std(train_data[eid]["X"][0])
        _, mean_y, std_y = _std(train_data[eid]["y"][0])
        for i in range(2):
            X = train_data[eid]["X"][i]
            K, T = X.shape[0], X.shape[1]
            X = (X - mean_X) / std_X
            if X.ndim == 2:
                X = X[:, :, None]
            X = np.concatenate([X.reshape(K, T, -1), np.ones((K, T, 1))], axis=2)
            train_data[eid]["X"][i] = X[:, sorted_idx]
            train_data[eid]["y"][i] = (train_data[eid]["y"][i] - mean_y) / std_y
            print(train_data[eid]["X"][i].shape, train_data[eid]["y"][i].shape)
        train_data[eid]["setup"].update(mean_X_Tv=mean_X, std_X_Tv=std_X, mean_y_TN=mean_y, std_y_TN=std_y)
    return train_data


def evaluate_session(pred, gt_held_out, threshold=1e-3):
    """src/train_rrr.py:198-224 (shared with utils.train_rrr)."""
    return evaluate_rrr_session(pred, gt_held_out, threshold)


def main(argv=None):
    args = get_args(argv)
    config = config_from_kwargs({"model": "include:{}".format(args.model_config)})
    config = update_config(args.train_config, config)
    config = update_config(args, config)
    set_seed(config.seed)
    sorted_idx = select_frames()                                  # train_rrr.py:48-49: first numpy draw after the seed

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    train_data = load_train_data(args.input_mod)
    eids = sorted(train_data.keys())
    ground_truth = {eid: train_data[eid]["y"][1] for eid in eids}
    my_eids = eids[rank::world]                                    # sessions are independent fits: shard them

    device_r0 = all(np.asarray(train_data[e]["X"][0]).dtype == np.uint8 for e in my_eids)
    if not device_r0:
        train_data = preprocess_host(train_data, my_eids, args.input_mod, sorted_idx)
    l2, n_comp = 100, 3
    # operand precision: None = the parity defaults (exact-operand mode for uint8 frames, 3 planes for float64 arrays);
    # VS_RRR_PLANES=1 selects the fastest, non-parity setting (DESIGN.md "RRR precision")
    planes = int(os.environ["VS_RRR_PLANES"]) if os.environ.get("VS_RRR_PLANES") else None
    print('start training')
    result, test_bps = {}, []
    for eid in my_eids:
        if device_r0:   # uint8 frames: smoothing, z-scoring, frame selection and packing on the device
            d = train_data[eid]
            entry = pack_session_from_frames(torch.from_numpy(d["X"][0]), d["y"][0], torch.from_numpy(d["X"][1]), d["y"][1],
                                             sorted_idx, n_comp, planes=planes, smooth_w=2.0)
            train_data[eid] = entry
        model, mse_val = train_model_main(train_data={eid: train_data[eid]}, l2=l2, n_comp=n_comp, save=True, planes=planes,
                                          model_fname='tmp' if world == 1 else f'tmp.rank{rank}')   # one writer per file
        print('finished training')
        print('eid:', eid)
        _, _, pred_orig = model.predict_y_fr(train_data, eid, 1)
        pred, bps_list, r2_list = evaluate_session(pred_orig.cpu().detach().numpy(), ground_truth[eid])
        co_bps = np.nanmean(bps_list)
        print(f"co-bps: {co_bps}")
        print(f"r2: {np.nanmean(r2_list)}")
        test_bps.append(co_bps)
        result[eid] = {'gt': ground_truth[eid], 'pred': pred, 'co_bps': bps_list, 'r2': r2_list, 'eid': eid}
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
        gathered = [None] * world
        dist.all_gather_object(gathered, result)
        result = {k: v for part in gathered for k, v in part.items()}
        test_bps = [np.nanmean(result[e]['co_bps']) for e in sorted(result)]
    if rank == 0:
        print(result.keys())
        for v in test_bps:
            print(f'{v:.5f}')
        print(f'mean bps:{np.mean(test_bps):.5f}')
        print(f"Total num of eid: {len(result.keys())}")
        np.save(f'{args.input_mod}_result.npy', result)
    return result


if __name__ == '__main__':
    main()
