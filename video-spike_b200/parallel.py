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
        self._gather = None

    # ---- device-driven mode: the collectives are stream-ordered (NCCL), so the whole step is still enqueued without a
    #      host read: dots (local shard) -> ONE all-gather of every rank's scalar block -> ordered combine -> update -> direction
    def _dev_dots(self, sp, n, g, g_prev, hist, f32, dev):
        ns = self._n_shared
        stride = hist.shape[1]
        if self._rank == 0 or self._world == 1 or ns == 0:
            self._dots_call(sp, n, vs.ptr(g), vs.ptr(g_prev), vs.ptr(hist), stride, f32, dev)
            return
        # rank != 0: the replicated prefix still gets its y = g - g_prev written (the history must be complete on every
        # rank), but its inner products are rank 0's to count: that pass's scalars are overwritten by the local pass
        esz = 4 if f32 else 8
        self._dots_call(sp, ns, vs.ptr(g), vs.ptr(g_prev), vs.ptr(hist), stride, f32, dev)
        if n > ns:
            self._dots_call(sp, n - ns, g.data_ptr() + 8 * ns, g_prev.data_ptr() + 8 * ns, hist.data_ptr() + esz * ns, stride, f32, dev)
        else:
            dev["out"].zero_()

    def _reduce_dev_scalars(self):
        if self._world == 1:
            return
        dev = self._dev
        out, dmax = dev["out"], dev["dmax"]
        k = out.numel()
        if self._gather is None or self._gather.device != out.device:
            self._gather = torch.empty((self._world, k + 1), dtype=torch.float64, device=out.device)
            self._mine = torch.empty(k + 1, dtype=torch.float64, device=out.device)
        self._mine[:k].copy_(out)
        self._mine[k:].copy_(dmax)
        dist.all_gather_into_tensor(self._gather.view(-1), self._mine, group=self._group)
        # sums in rank order (bit-identical on every rank whatever the collective's internal order); |g|_inf and max|t*d| by max.
        # entries past 8 + 6m hold stale values on every rank: they are never read
        torch.sum(self._gather[:, :k], dim=0, out=out)
        out[2:3].copy_(self._gather[:, 2].max().reshape(1))
        dmax.copy_(self._gather[:, k].max().reshape(1))

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


def train_joint_model(model, train_data_local, model_fname="tmp", save=False, group=None, history_dtype=None, device_driven=None):
    """`train_model` (rrr.py:164-190) for the sharded joint model: one ShardedLBFGS.step over all ranks' sessions,
    validation SSE summed over ranks.  Returns (model, {"mses_val": local dict, "mse_val_mean": global})."""
    shared = [model.model["V"]]
    local = [p for k_, p in model.model.items() if k_ != "V"]
    if history_dtype is None:
        history_dtype = torch.float32 if model.planes == 1 else torch.float64
    if device_driven is None:
        device_driven = os.environ.get("VS_LBFGS_DEVICE", "1") != "0" and model.model["V"].is_cuda
    compact = bool(device_driven and history_dtype == torch.float64 and os.environ.get("VS_LBFGS_COMPACT", "1") != "0")
    optimizer = ShardedLBFGS(shared, local, group=group, history_dtype=history_dtype, device_driven=device_driven, compact=compact)
    optimizer.closure_overwrites_grads = True

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


# =====================================================================================================
# ONE session over n GPUs (third row): trials are sharded, the parameters are replicated.
def _combine_stats(mean, std, k_local, group=None):
    """Global mean / clipped population std over all ranks' trials from the local ones (Chan's parallel formula,
    float64): what src/utils/utils.py:107-112 computes over the whole train split."""
    k = torch.tensor([float(k_local)], dtype=torch.float64, device=mean.device)
    s1 = mean * k_local
    s2 = (std * std + mean * mean) * k_local          # sum x^2 = K (var + mean^2)
    for t in (k, s1, s2):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    gmean = s1 / k
    gvar = (s2 / k - gmean * gmean).clamp_min(0.0)
    return gmean, gvar.sqrt().clamp_min(1e-8), int(k.item())


def pack_trial_shard(frames_train, counts_train, frames_test, counts_test, sorted_idx, n_comp, planes=1, smooth_w=2.0,
                     device=None, operand=None, group=None, mode=None):
    """`model.rrr.pack_session_from_frames` for THIS rank's trials of a session whose trials are spread over the ranks:
    the z-score statistics (frames and smoothed counts) are those of the WHOLE train split (three small all-reduces),
    so every rank's operands are exactly the rows it would own in the single-GPU pack.
    mode "exact" / "dense": the exact-operand layouts of pack_session_from_frames (its loader, with the combined statistics);
    mode None / "classic": `planes` planes of the z-score."""
    from model.rrr import _PackedSplit, _op_dtype, operand_format, pack_session_from_frames
    vs.require_b200()
    if mode not in (None, "classic"):
        def hook(mean, sd, my, sy, K):
            if world()[1] > 1:
                mean, sd, _ = _combine_stats(mean, sd, K, group)
                my, sy, _ = _combine_stats(my, sy, K, group)
            return mean, sd, my, sy
        return pack_session_from_frames(frames_train, counts_train, frames_test, counts_test, sorted_idx, n_comp, smooth_w=smooth_w,
                                        device=device, mode=mode, stats_hook=hook)
    device = device or torch.device("cuda")
    st = vs.stream()
    idx = torch.as_tensor(np.asarray(sorted_idx), dtype=torch.int32).to(device)
    T = int(idx.numel())
    fmt = operand_format(planes, operand)
    splits, ys = [], []
    mean = sd = my = sy = None
    for which, (fr, cnt) in enumerate(((frames_train, counts_train), (frames_test, counts_test))):
        fr = fr.to(device, non_blocking=True).reshape(fr.shape[0], fr.shape[1], -1).contiguous()
        cnt = torch.as_tensor(cnt).to(device, non_blocking=True).float().contiguous()
        K, Tf, F = fr.shape
        N = cnt.shape[2]
        if which == 0:
            mean = torch.empty(Tf * F, dtype=torch.float64, device=device)
            sd = torch.empty_like(mean)
            vs.check(vs.lib.vs_rrr_colstats(vs.ptr(fr), K, Tf * F, vs.ptr(mean), vs.ptr(sd), st))
            sm = torch.empty_like(cnt)
            vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, N, float(smooth_w), None, None, vs.ptr(sm), st))
            my = torch.empty(T * N, dtype=torch.float64, device=device)
            sy = torch.empty_like(my)
            vs.check(vs.lib.vs_colstats_f32(vs.ptr(sm), K, T * N, vs.ptr(my), vs.ptr(sy), st))
            del sm
            if world()[1] > 1:
                # the kernels clip std at 1e-8 before returning it: a locally constant column reports 1e-8, whose square
                # (1e-16) is far below float64 noise of the combined second moment -- harmless in the combination
                mean, sd, _ = _combine_stats(mean, sd, K, group)
                my, sy, _ = _combine_stats(my, sy, K, group)
        d = vs.RrrDims(K, T, F, N, n_comp, planes, vs.lib.vs_rrr_ldc(F), vs.lib.vs_rrr_ldr(K, T), fmt)
        Xa = torch.empty((planes, K * T, d.ldc), dtype=_op_dtype(fmt), device=device)
        Xb = torch.empty((planes, F, d.ldr), dtype=_op_dtype(fmt), device=device)
        xl = torch.empty(K * T, dtype=torch.float32, device=device)
        overflow = torch.zeros(1, dtype=torch.int32, device=device)
        vs.check(vs.lib.vs_rrr_pack_u8(vs.ptr(fr), Tf, vs.ptr(idx), vs.ptr(mean), vs.ptr(sd), d, vs.ptr(Xa), vs.ptr(Xb),
                                       vs.ptr(xl), vs.ptr(overflow), st))
        y = torch.empty_like(cnt)
        vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, N, float(smooth_w), vs.ptr(my), vs.ptr(sy), vs.ptr(y), st))
        splits.append(_PackedSplit.from_device(d, Xa, Xb, xl, y, overflow))
        ys.append(y)
        del fr
    return {"X": splits, "y": ys,
            "setup": {"mean_X_Tv": mean, "std_X_Tv": sd, "mean_y_TN": my.reshape(T, -1), "std_y_TN": sy.reshape(T, -1)}}


def build_trial_sharded_model(entry_local, l2, n_comp, eid="session", planes=None, operand=None, group=None):
    """The replicated RRRGD model of a trial-sharded session: l2 / world (the penalty is linear in l2, so the ranks' losses
    and gradients SUM to the single-GPU ones), b = mean over ALL trials of the session (rrr.py:47, local sums all-reduced)."""
    from model.rrr import RRRGD, get_device
    rank, ws = world()
    td = {eid: entry_local}
    if planes is None:
        planes = entry_local["X"][0].dims.planes                   # the layout the shard was packed in
    device = get_device()
    # the init stream (8 M normals on the host cores) is drawn by rank 0 only and broadcast: every rank drawing the same
    # stream made the ranks compete for the host cores (8 ranks: 4x slower than one)
    model = RRRGD(td, n_comp, l2=l2 / ws, planes=planes, operand=operand, device=device, draw=(rank == 0 or ws == 1))
    yl = entry_local["y"][0]
    kk = torch.tensor([float(yl.shape[0])], dtype=torch.float64, device=yl.device)
    bsum = yl.double().sum(0).T.unsqueeze(1).contiguous()
    if ws > 1:
        dist.all_reduce(kk, group=group)
        dist.all_reduce(bsum, group=group)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        with torch.no_grad():
            dist.broadcast(model.model[f"{eid}_U"].data, src=src, group=group)
            dist.broadcast(model.model["V"].data, src=src, group=group)
    with torch.no_grad():
        model.model[f"{eid}_b"].copy_(bsum / kk)
    return model


def fit_trial_sharded(model, entry_local, eid="session", group=None, history_dtype=None):
    """One L-BFGS step (rrr.py:164-190) of a trial-sharded session + the validation SSE over all ranks' validation trials.
    Parameters and the whole L-BFGS state are replicated; each closure evaluation runs on the local trials followed by
    ONE all-reduce of [flat gradient] and one of the loss; every rank then takes the identical L-BFGS step."""
    rank, ws = world()
    td = {eid: entry_local}
    optimizer = model.make_optimizer(history_dtype=history_dtype)

    def closure():
        optimizer.zero_grad()
        model.train()
        loss = model.loss_and_grad(td, 0)
        if ws > 1:
            fl = optimizer._flat
            g = fl["g"][fl["cur"]]                                  # every p.grad is a view of this one buffer
            buf = loss.reshape(1).clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            loss = buf[0]
        return loss

    optimizer.step(closure)
    model.eval()
    mses_val = model.compute_MSE_RRRGD(td, 1)
    total = torch.sum(mses_val[eid]).clone()
    per_neuron = mses_val[eid].clone()
    if ws > 1:
        dist.all_reduce(total, group=group)
        dist.all_reduce(per_neuron, group=group)
    return model, {"mses_val": {eid: per_neuron}, "mse_val_mean": total}


def train_trial_sharded(entry_local, l2, n_comp, eid="session", planes=None, operand=None, group=None, history_dtype=None):
    """`train_model_main` (rrr.py:192-202) for ONE session whose trials are sharded over the ranks
    (build_trial_sharded_model + fit_trial_sharded)."""
    model = build_trial_sharded_model(entry_local, l2, n_comp, eid, planes, operand, group)
    return fit_trial_sharded(model, entry_local, eid, group, history_dtype)
