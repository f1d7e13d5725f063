"""Per-trial shard loader: the role of the reference's src/loader/base.py (webdataset tar shards -> dict batches
{'ap' (B,100,N) float32, <input mods>, 'eid' [B]}), without webdataset / mp4 decoding (neither `webdataset` nor `av`
is in this image).  Supported trial containers, one trial per file like the reference's
`<eid>_<trial>.tar` (src/prepare_data.py:210-233):

  *.tar   members `<key>.ap.pyd`, `<key>.<mod>.pyd` (pickled numpy, webdataset's "pyd") and `<key>.video.npy` /
          `<key>.whisker-video.npy`: raw uint8 (T,H,W) or (T,H,W,C) frames -- the one-off transcode of the mp4 member
  *.npz   arrays `ap`, `video`, ... with the same meaning

Video stays **uint8** (B,T,1,H,W): the `.float()` of src/loader/base.py:39 is folded into the first-layer kernels, which
read the bytes directly (values 0..255, never rescaled: SURVEY A1).  Every other modality is cast to float32 like the
reference.  `synthetic:` data_dirs generate seeded trials of the loader's shape for benchmarking.
"""
from __future__ import annotations

import io
import os
import pickle
import random
import tarfile

import numpy as np
import torch

VIDEO_MODS = ("video", "whisker-video")


def _process(mod, value):
    """src/loader/base.py:43-95 (process_modalities): first channel of the video, (T,1,H,W); the rest from_numpy."""
    if mod in VIDEO_MODS:
        v = torch.as_tensor(np.asarray(value))
        if v.ndim == 4:                      # (T,H,W,C) as decoded by torchvision: grayscale, take channel 0
            v = v[:, :, :, 0]
        return v.unsqueeze(1).contiguous()   # uint8 (T,1,H,W)
    return torch.as_tensor(np.asarray(value)).float()


def read_trial(path, modalities):
    out = {}
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            for k in z.files:
                if k in modalities:
                    out[k] = _process(k, z[k])
    else:
        with tarfile.open(path) as tf:
            for m in tf.getmembers():
                parts = os.path.basename(m.name).split(".")
                if len(parts) < 3 or parts[-2] not in modalities:
                    continue
                raw = tf.extractfile(m).read()
                if parts[-1] == "pyd":
                    out[parts[-2]] = _process(parts[-2], pickle.loads(raw))
                elif parts[-1] == "npy":
                    out[parts[-2]] = _process(parts[-2], np.load(io.BytesIO(raw), allow_pickle=False))
    out["eid"] = os.path.basename(path).split("_")[0]          # src/loader/base.py:40
    return out


class TrialDataset(torch.utils.data.Dataset):
    def __init__(self, config, files, mode="train"):
        self.files = list(files)
        self.mods = set(config.data.modalities.keys())
        if mode == "train":
            random.Random(config.seed).shuffle(self.files)    # the reference shuffles with Random(config.seed)

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        return read_trial(self.files[i], self.mods)


class SyntheticTrials(torch.utils.data.Dataset):
    """`synthetic:n=64,h=32,w=32,neurons=20[,seed=0]` -- sparse uint8 frames (the value distribution on which the
    reference itself trains stably) + Poisson(0.3) counts, one fixed eid."""

    def __init__(self, spec, names):
        kv = dict(item.split("=") for item in spec.split(",") if item)
        self.h, self.w, self.N = int(kv.get("h", 32)), int(kv.get("w", 32)), int(kv.get("neurons", 20))
        self.seed = int(kv.get("seed", 0))
        self.names = list(names)

    def __len__(self):
        return len(self.names)

    def __getitem__(self, i):
        name = self.names[i]
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + int(name.split("_")[-1].split(".")[0]))
        mask = torch.rand((120, 1, self.h, self.w), generator=g) < 0.01
        video = (torch.randint(1, 9, (120, 1, self.h, self.w), generator=g) * mask).to(torch.uint8)
        ap = torch.poisson(torch.full((100, self.N), 0.3), generator=g)
        return {"video": video, "ap": ap, "eid": name.split("_")[0]}


def collate(samples):
    out = {}
    for k in samples[0]:
        vals = [s[k] for s in samples]
        out[k] = torch.stack(vals) if isinstance(vals[0], torch.Tensor) else vals
    return out
