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
