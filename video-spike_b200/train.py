"""`Linear` driver: drop-in for the reference's src/train.py (same CLI, config handling, optimizer/scheduler recipe,
trainer wiring), with FusedAdamW + the fused train step doing the arithmetic on the B200.

    python train.py --model_config config/model/linear_video.yaml --train_config config/train/linear_video.yaml --eid <eid>

`config.dirs.data_dir` may be a directory of per-trial shards (loader/base.py) or `synthetic:n=64,h=32,w=32,neurons=20`.
"""
from __future__ import annotations

import os
import random
import sys

import torch
from torch.optim.lr_scheduler import OneCycleLR

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from loader.make import make_loader  # noqa: E402
from optim import FusedAdamW  # noqa: E402
from trainer.make import make_base_trainer  # noqa: E402
from utils.accel import Accelerator  # noqa: E402
from utils.config_utils import config_from_kwargs, update_config  # noqa: E402
from utils.dataset_utils import get_eids_from_filenames, get_metadata_from_loader, split_dataset  # noqa: E402
from utils.utils import NAME2MODEL, get_args, set_seed  # noqa: E402


def synthetic_split(spec, eid, train_ratio=0.8, val_ratio=0.1):
    """split_dataset (src/utils/dataset_utils.py:50-88) over generated trial names: same shuffle, same cut points."""
    kv = dict(item.split("=") for item in spec.split(",") if item)
    files = [f"{eid}_{i}.tar" for i in range(int(kv.get("n", 64)))]
    random.shuffle(files)
    c1, c2 = int(train_ratio * len(files)), int((train_ratio + val_ratio) * len(files))
    tr, va, te = files[:c1], files[c1:c2], files[c2:]
    return {"train": tr, "val": va, "test": te,
            "eid": {"train": get_eids_from_filenames(tr), "val": get_eids_from_filenames(va), "test": get_eids_from_filenames(te)}}


def main(argv=None):
    args = get_args(argv)
    config = config_from_kwargs({"model": "include:{}".format(args.model_config)})
    config = update_config(args.train_config, config)
    config = update_config(args, config)
    set_seed(config.seed)
    data_dir = str(config.dirs.data_dir)
    if data_dir.startswith("synthetic:"):
        dataset_split_dict = synthetic_split(data_dir[len("synthetic:"):], args.eid)
    else:
        dataset_split_dict = split_dataset(data_dir, eid=args.eid)
    train_dataloader, val_dataloader, test_dataloader = make_loader(config, dataset_split_dict)
    meta_data = get_metadata_from_loader(train_dataloader, config)
    print(f"meta_data: {meta_data}")
    model_class = NAME2MODEL[config.model.model_class]
    config['model']['encoder']['input_dim'] = meta_data['input_dim']       # YAML sizes are overwritten
    config['model']['decoder']['output_dim'] = meta_data['output_dim']
    model = model_class(config.model)
    optimizer = FusedAdamW(model.parameters(), lr=config.optimizer.lr, weight_decay=config.optimizer.wd, eps=config.optimizer.eps)
    lr_scheduler = OneCycleLR(optimizer=optimizer,
                              total_steps=len(dataset_split_dict['train']) // config.training.train_batch_size * config.training.num_epochs,
                              max_lr=config.optimizer.lr, pct_start=config.optimizer.warmup_pct, div_factor=config.optimizer.div_factor)
    criterion = torch.nn.PoissonNLLLoss(reduction="none", log_input=True)
    accelerator = Accelerator()
    model, optimizer, lr_scheduler = accelerator.prepare(model, optimizer, lr_scheduler)
    trainer = make_base_trainer(model=model, optimizer=optimizer, train_dataloader=train_dataloader, eval_dataloader=val_dataloader,
                                test_dataloader=test_dataloader, log_dir=config.dirs.log_dir, accelerator=accelerator,
                                lr_scheduler=lr_scheduler, config=config, criterion=criterion, dataset_split_dict=dataset_split_dict,
                                eid=args.eid)
    trainer.train()
    return trainer


if __name__ == '__main__':
    main()
