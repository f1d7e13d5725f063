"""Supervised trainer: drop-in for the reference's src/trainer/base.py.

Same constructor kwargs (log_dir, accelerator, lr_scheduler, config, criterion, dataset_split_dict,
eid), same methods and return dictionaries (`train`, `train_epoch -> {"train_loss","lr"}`,
`eval_epoch -> {"eval_gt","eval_preds","eval_res"}`, `test_model`, `computer_loss`,
`_forward_model_outputs`, `save_model`), same log directory layout and whole-module checkpoints.

The step body (base.py:147-154) has two routes:
  * fused  -- model is model.linear.Linear, optimizer is optim.FusedAdamW and the criterion is
              PoissonNLLLoss(log_input=True, reduction="none"): ONE vs_mlp_train_step call per batch
              (uint8 frames in, every parameter updated on the device, loss read back once per epoch
              instead of a host sync per step).
  * generic -- any other combination: model(...) / criterion / accelerator.backward / optimizer.step
              exactly as the reference sequences them (the model's kernels still do the arithmetic).
Plotting (matplotlib) and wandb are optional and skipped when the packages are absent.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from tqdm import tqdm

from utils.utils import metrics_list, move_batch_to_device

try:  # optional, exactly as optional as config.wandb.use makes it
    import wandb
except Exception:  # pragma: no cover
    wandb = None


def _get_input_modailities(config):
    """Input modalities in YAML key order (base.py:8-14; the order fixes the concat layout)."""
    return [m for m in config.data.modalities.keys() if config.data.modalities[m]["input"]]


class BaseTrainer():
    def __init__(self, model, train_dataloader, eval_dataloader, test_dataloader, optimizer, **kwargs):
        self.model = model
        self.train_dataloader = train_dataloader
        self.eval_dataloader = eval_dataloader
        self.test_dataloader = test_dataloader
        self.optimizer = optimizer
        self.criterion = kwargs.get("criterion", None)
        self.log_dir = kwargs.get("log_dir", None)
        self.accelerator = kwargs.get("accelerator", None)
        self.lr_scheduler = kwargs.get("lr_scheduler", None)
        self.config = kwargs.get("config", None)
        self.dataset_split_dict = kwargs.get("dataset_split_dict", None)
        self.eid = kwargs.get("eid", None)
        self.model_class = self.config.model.model_class
        self.metrics = ['bps', 'rsquared']
        self.session_active_neurons = {}
        self.input_mods = _get_input_modailities(self.config)
        self._create_log_dir()

    # ------------------------------------------------------------------ bookkeeping
    def _use_wandb(self):
        return bool(self.config.wandb.use) and wandb is not None

    def _create_log_dir(self):
        mods = "_".join(self.input_mods)
        name = self.model.__class__.__name__
        self.log_dir = os.path.join(self.log_dir, self.eid[:5], mods, name)
        os.makedirs(self.log_dir, exist_ok=True)
        if self._use_wandb():
            wandb.init(project=self.config.wandb.project, name="{}_{}_{}".format(self.eid[:5], mods, name), config=self.config)

    def _unwrapped(self):
        """(model, optimizer) without the wrappers `accelerator.prepare` may add (DDP / AcceleratedOptimizer)."""
        model = self.model
        unwrap = getattr(self.accelerator, "unwrap_model", None)
        if unwrap is not None:
            model = unwrap(model)
        model = getattr(model, "module", model)
        return model, getattr(self.optimizer, "optimizer", self.optimizer)

    def _is_main(self):
        return bool(getattr(self.accelerator, "is_main_process", True))

    def _fused_route(self):
        from model.linear import Linear
        from optim import FusedAdamW
        c = self.criterion
        model, optimizer = self._unwrapped()
        if model is not self.model:
            # a wrapped (DDP) model synchronises gradients through autograd hooks: the fused step and the factored
            # first-layer gradient bypass them, so a wrapped model always takes the generic route with plain gradients
            for p in model.parameters():
                p._vs_defer_ok = False
            return False
        return (isinstance(model, Linear) and isinstance(optimizer, FusedAdamW)
                and isinstance(c, torch.nn.PoissonNLLLoss) and c.log_input and c.reduction == "none" and not c.full)

    # ------------------------------------------------------------------ forward / loss
    def _gather_inputs(self, batch):
        batch = move_batch_to_device(batch, self.accelerator.device)
        if self.config.model.model_class == "Linear":
            parts = [batch[mod].flatten(1) for mod in self.input_mods]
            if len(parts) == 1:
                return parts[0]               # uint8 frames go to the kernels untouched
            return torch.cat([p.float() for p in parts], dim=-1)
        return batch['video']

    def _forward_model_outputs(self, batch):
        return self.model(self._gather_inputs(batch))

    def computer_loss(self, outputs, batch):
        return self.criterion(outputs, batch['ap']).mean()

    # ------------------------------------------------------------------ training
    def train_epoch(self):
        losses = []
        self.model.train()
        fused = self._fused_route()
        for batch in tqdm(self.train_dataloader):
            if fused:
                inputs = self._gather_inputs(batch)
                loss = self.model.fused_train_step(inputs, batch['ap'], self._unwrapped()[1])
                self.lr_scheduler.step()
                losses.append(loss)               # device scalar: no host sync inside the loop
            else:
                outputs = self._forward_model_outputs(batch)
                loss = self.computer_loss(outputs, batch)
                self.accelerator.backward(loss)
                self.optimizer.step()
                self.lr_scheduler.step()
                self.optimizer.zero_grad()
                losses.append(loss.detach())
        vals = torch.stack([l.double().reshape(()) for l in losses]).cpu().numpy() if losses else np.array([np.nan])
        return {"train_loss": round(float(np.mean(vals)), 5), "lr": self.optimizer.param_groups[0]['lr']}

    def train(self):
        best_eval_loss = torch.tensor(float('inf'))
        best_eval_bps = -torch.tensor(float('inf'))
        print("start training")
        epoch = 0
        for epoch in range(self.config.training.num_epochs):
            train_res = self.train_epoch()
            eval_res = self.eval_epoch()
            print(f"epoch: {epoch} train loss: {train_res['train_loss']}")
            if eval_res:
                if eval_res['eval_res']['eval_bps'] > best_eval_bps:
                    best_eval_bps = eval_res['eval_res']['eval_bps']
                    best_eval_loss = eval_res['eval_res']['eval_loss']
                    print(f"epoch: {epoch} best eval_bps: {best_eval_bps}")
                    self.save_model(name="best", epoch=epoch)
                    if self._use_wandb():
                        wandb.log({"best_eval_bps_epoch": epoch})
                    print(f"best_epoch: {epoch}, best_eval_bps: {best_eval_bps}")
                log = {**train_res, **eval_res['eval_res']}
                wandb.log(log) if self._use_wandb() else print(log)
        self.save_model(name="last", epoch=epoch)
        test_res = self.test_model()
        if test_res:
            log = {**test_res['test_res'], "best_eval_loss": best_eval_loss, "best_eval_bps": best_eval_bps}
            np.save(os.path.join(self.log_dir, "test_results.npy"), test_res)
            wandb.log(log) if self._use_wandb() else print(log)

    # ------------------------------------------------------------------ evaluation
    def _run_split(self, loader, eids, phase):
        """base.py:161-206 / 209-256: no-grad forward per batch, exp(pred), per-session metrics."""
        losses = []
        per_eid = {eid: {'gt': [], 'preds': []} for eid in eids}
        gt, preds = {}, {}
        metrics_results = {k: [] for k in self.metrics}
        if loader is not None:
            for batch in loader:
                outputs = self._forward_model_outputs(batch)
                losses.append(self.computer_loss(outputs, batch).item())
                eid = batch['eid'][0]                      # one session per eval batch
                per_eid[eid]['gt'].append(batch['ap'])
                per_eid[eid]['preds'].append(outputs)
            for idx, eid in enumerate(eids):
                g = torch.cat(per_eid[eid]['gt'], dim=0)
                p = torch.exp(torch.cat(per_eid[eid]['preds'], dim=0))
                gt[idx], preds[idx] = g, p
                res = metrics_list(gt=g.transpose(-1, 0), pred=p.transpose(-1, 0), metrics=self.metrics,
                                   device=self.accelerator.device)
                for k, v in res.items():
                    metrics_results[k].append(v)
        summary = {f"{phase}_{k}": round(np.mean(v), 5) for k, v in metrics_results.items()}
        return {f"{phase}_gt": gt, f"{phase}_preds": preds,
                f"{phase}_res": {f"{phase}_loss": round(np.mean(losses), 5), **summary}}

    @torch.no_grad()
    def eval_epoch(self):
        self.model.eval()
        return self._run_split(self.eval_dataloader, self.dataset_split_dict['eid']['val'], "eval")

    @torch.no_grad()
    def test_model(self):
        # the best checkpoint is a pickled module (base.py:212); rank 0 wrote it
        wait = getattr(self.accelerator, "wait_for_everyone", None)
        if wait is not None:
            wait()
        self.model = torch.load(os.path.join(self.log_dir, "model_best.pt"), weights_only=False)['model']
        self.model.eval()
        return self._run_split(self.test_dataloader, self.dataset_split_dict['eid']['test'], "test")

    def save_model(self, name="last", epoch=0):
        if not self._is_main():        # every rank trains the same replica (the reference never shards its loaders): one writer
            return
        print(f"saving model: {name} to {self.log_dir}")
        torch.save({"model": self._unwrapped()[0], "epoch": epoch}, os.path.join(self.log_dir, f"model_{name}.pt"))
