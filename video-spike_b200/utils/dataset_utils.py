"""Dataset bookkeeping with the behaviour of the reference's src/utils/dataset_utils.py:
80/10/10 split of per-trial tar shards (:50-88), eid extraction (:90-97), and the first-batch probe
that fixes the model's input/output sizes (:99-119).
"""
from __future__ import annotations

import os
import random

import torch


def split_dataset(data_dir, eid, train_ratio=0.8, val_ratio=0.1, test_ratio=0.1):
    """List *.tar shards of the session(s) `eid` (str or list), shuffle with the global `random`
    stream and cut at int(0.8 n) and int(0.9 n) -- same order of operations and same cut points as
    the reference, so the split is identical under the same seed and directory listing."""
    wanted = [eid] if isinstance(eid, str) else list(eid)
    files = [os.path.join(data_dir, f) for f in os.listdir(data_dir) if f.endswith(".tar")]
    files = [f for f in files if any(e in f for e in wanted)]
    print(f"Found {len(files)} files for EID: {wanted}")
    random.shuffle(files)
    cut1 = int(train_ratio * len(files))
    cut2 = int((train_ratio + val_ratio) * len(files))
    train, val, test = files[:cut1], files[cut1:cut2], files[cut2:]
    return {
        "train": train, "val": val, "test": test,
        "eid": {"train": get_eids_from_filenames(train), "val": get_eids_from_filenames(val),
                "test": get_eids_from_filenames(test)},
    }


def get_eids_from_filenames(filenames):
    return list({os.path.basename(f).split("_")[0] for f in filenames})


def get_metadata_from_loader(data_loader, config):
    """Pull ONE batch; input_dim = total flattened width of the input modalities (YAML key order),
    output_dim = time bins x neurons of `ap`."""
    batch = next(iter(data_loader))
    input_mods = [m for m in config.data.modalities.keys() if config.data.modalities[m]["input"]]
    width = sum(int(batch[m].flatten(1).shape[1]) for m in input_mods)
    ap = batch["ap"]
    return {"num_neurons": ap.shape[2], "input_dim": width, "input_mods": input_mods,
            "output_dim": ap.shape[1] * ap.shape[2]}


def load_video_index(timestamps, intervals, fps):
    """Frame window of every trial (src/utils/ibl_data_utils.py:958-967): `int(fps * interval_len)` frames
    (120 at 60 Hz x 2 s) starting at the first frame whose timestamp is >= the interval start
    (`np.searchsorted`, side='left').  Integer work, bit-exact by construction; raises like the reference
    when a trial holds more than 10 frames too many/few."""
    import numpy as np
    ts = np.asarray(timestamps)
    intervals = np.asarray(intervals)
    reg_frame_num = int(fps * (intervals[0, 1] - intervals[0, 0]))
    out = np.empty((len(intervals), reg_frame_num), dtype=np.int64)
    for i, (t0, t1) in enumerate(intervals):
        n_in = int(np.count_nonzero((ts > t0) & (ts < t1)))
        if abs(n_in - reg_frame_num) > 10:
            raise ValueError(f"Number of frames in the video does not match the expected number of frames {reg_frame_num}. Bias > 10")
        start = int(np.searchsorted(ts, t0))
        out[i] = np.arange(start, start + reg_frame_num)
    return out


def gather_trial_windows(session_frames_u8, start_idx, frames_per_trial=120):
    """(n_frames, ...) uint8 session video on the DEVICE + per-trial start indices -> (n_trials, 120, ...) uint8:
    the windows `load_video_index` defines, cut by the vs_gather_windows kernel (bit-exact byte copy)."""
    import vsb200 as vs
    fr = session_frames_u8
    starts = torch.as_tensor(start_idx, dtype=torch.int64, device=fr.device).contiguous()
    row = int(fr[0].numel())
    out = torch.empty((starts.numel(), frames_per_trial) + tuple(fr.shape[1:]), dtype=torch.uint8, device=fr.device)
    if starts.numel():
        vs.check(vs.lib.vs_gather_windows(vs.ptr(fr.contiguous()), fr.shape[0], row, vs.ptr(starts), starts.numel(),
                                          frames_per_trial, vs.ptr(out), vs.stream()))
    return out
