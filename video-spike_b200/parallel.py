"""Multi-GPU plumbing for the RRR path.  One process per GPU (torchrun / accelerate launch env contract);
`torch.distributed` (NCCL over NVLink on the box, gloo in the CPU tests) carries the only exchange the path has.

* Independent per-session fits (what src/train_rrr.py:179-187 actually runs): `shard_sessions` -- no collective.
* ONE joint model over all sessions with a shared V (what `RRRGD` supports, src/model/rrr.py:37-49):
  every rank holds U_s, b_s, X_s, y_s of ITS sessions plus a replica of V.  Per closure evaluation one
  all-reduce(sum) of [dV (r*T values), loss]; per L-BFGS iteration one all-reduce(sum) of the inner products and
  one all-reduce(max) of [|g|_inf, max|t*d|] (`ShardedLBFGS`).  The optimiser state that depends on those scalars
  (Gram matrix, coefficients, step length, termination) is then bit-identical on every rank, so V stays replicated
  without ever being broadcast.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

import vsb200 as vs
from optim import FusedLBFGS


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_sessions(eids, rank=None, world_size=None, cost=None):
    """Sessions of this rank.  Without `cost`: round-robin over the sorted eids (train_rrr.py:112 sorts them).
    With `cost` {eid: work estimate, e.g. K*C*N}: longest-processing-time-first bin packing, deterministic."""
    if rank is None or world_size is None:
        rank, world_size = world()
    eids = sorted(eids)
    if cost is None:
        return eids[rank::world_size]
    load = [0.0] * world_size
    mine = []
    for e in sorted(eids, key=lambda e: (-float(cost[e]), e)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        load[r] += float(cost[e])
        if r == rank:
            mine.append(e)
    return sorted(mine)


class ShardedLBFGS(FusedLBFGS):
    """FusedLBFGS over a parameter vector that is partly replicated (`shared`, e.g. V) and partly rank-local
    (`local`, e.g. this rank's U_s, b_s).  Mathematically it is torch.optim.LBFGS on the concatenation of all ranks'
    local parameters and ONE copy of the shared ones: shared elements enter every inner product once (rank 0 counts
    them), and their gradient must already be the global one (the closure all-reduces it)."""

    def __init__(self, shared, local, group=None, **kw):
        shared, local = list(shared), list(local)
        super().__init__(shared + local, **kw)                 # shared block first in the flat vector
        self._n_shared = sum(p.numel() for p in shared)
        self._group = group
        self._rank, self._world = world()

    def _dots(self, g, g_prev, s_slot, y_slot, loss_t):
        k = 8 + 6 * len(self._pairs)
        n, ns = self._flat["n"], self._n_shared
        if self._rank == 0 or self._world == 1 or ns == 0:
            self._pass_dots(0, n, g, g_prev, s_slot, y_slot, self._scal)
        else:
            # the shared prefix still needs its y = g - g_prev (history must be complete on every rank) but its inner
            # products are rank 0's to count: run it into a scratch output, then the local part for real
            if y_slot is not None:
                scratch = torch.empty_like(self._scal)
                self._pass_dots(0, ns, g, g_prev, s_slot, y_slot, scratch)
            if n > ns:
                self._pass_dots(ns, n, g, g_prev, s_slot, y_slot, self._scal)
            else:
                self._scal[:k].zero_()
        self._reduce_scalars(k)
        self._scal[-1:].copy_(torch.as_tensor(loss_t).detach().reshape(1))
        host = self._scal.cpu().numpy()
        return host[:k], float(host[-2]), float(host[-1])

    def _reduce_scalars(self, k):
        if self._world == 1:
            return
        mx = torch.stack([self._scal[2], self._scal[-2]])      # |g|_inf and max|t*d| combine by max
        dist.all_reduce(self._scal[:k], op=dist.ReduceOp.SUM, group=self._group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self._group)
        self._scal[2] = mx[0]
        self._scal[-2] = mx[1]


def joint_loss_and_grad(model, data, k=0, group=None):
    """Closure body of src/model/rrr.py:165-175 for a session-sharded joint model: local sessions' loss and
    gradients, then ONE all-reduce of [dV, loss] so every rank holds the global V gradient and the global loss."""
    loss = model.loss_and_grad(data, k)
    rank, ws = world()
    if ws > 1:
        V = model.model["V"]
        buf = torch.cat([V.grad.reshape(-1), loss.reshape(1)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        V.grad.copy_(buf[:-1].view_as(V.grad))
        loss = buf[-1]
    return loss


def train_joint_model(model, train_data_local, model_fname="tmp", save=False, group=None, history_dtype=None):
    """`train_model` (rrr.py:164-190) for the sharded joint model: one ShardedLBFGS.step over all ranks' sessions,
    validation SSE summed over ranks.  Returns (model, {"mses_val": local dict, "mse_val_mean": global})."""
    shared = [model.model["V"]]
    local = [p for k_, p in model.model.items() if k_ != "V"]
    if history_dtype is None:
        history_dtype = torch.float32 if model.planes == 1 else torch.float64
    optimizer = ShardedLBFGS(shared, local, group=group, history_dtype=history_dtype)

    def closure():
        optimizer.zero_grad()
        model.train()
        return joint_loss_and_grad(model, train_data_local, 0, group)

    optimizer.step(closure)
    model.eval()
    mses_val = model.compute_MSE_RRRGD(train_data_local, 1)
    total = torch.sum(torch.cat([mses_val[k_] for k_ in mses_val])) if mses_val else torch.zeros((), dtype=torch.float64, device=model.model["V"].device)
    if world()[1] > 1:
        total = total.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    if save and world()[0] == 0:
        torch.save({"RRRGD_model": model.state_dict(), "optimizer": {}}, model_fname)
    return model, {"mses_val": mses_val, "mse_val_mean": total}
