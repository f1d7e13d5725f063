"""make_loader(config, dataset_split_dict) with the reference's signature (src/loader/make.py): train / val / test
DataLoaders over per-trial shards.  Pinned host batches so the H2D copy of the uint8 frames is asynchronous."""
from __future__ import annotations

import torch

from loader.base import SyntheticTrials, TrialDataset, collate


def make_loader(config, dataset_split_dict):
    data_dir = str(config.dirs.data_dir)
    loaders = []
    for mode in ("train", "val", "test"):
        files = dataset_split_dict[mode]
        if data_dir.startswith("synthetic:"):
            ds = SyntheticTrials(data_dir[len("synthetic:"):], files)
        else:
            ds = TrialDataset(config, files, mode)
        bs = config.training.train_batch_size if mode == "train" else config.training.test_batch_size
        loaders.append(torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=False, num_workers=int(config.training.num_workers),
                                                   collate_fn=collate, pin_memory=torch.cuda.is_available()))
    return tuple(loaders)
