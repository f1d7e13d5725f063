"""CPU: the N>1 host logic over gloo (world_size 2) and the optimiser's host logic on a torch backend.
The vector passes are the torch stand-ins of tests/cpu_lbfgs_backend.py; everything else is the product code
(optim.FusedLBFGS, parallel.ShardedLBFGS, parallel.shard_sessions)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.cpu_lbfgs_backend import TorchPasses


def _objective(seed=0, n_shared=12, n_local=(40, 25)):
    """f(shared, local_0, local_1) = sum_r f_r(shared, local_r): smooth, ill-conditioned, non-quadratic."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for nl in n_local:
        n = n_shared + nl
        A = torch.randn(n, n, generator=g, dtype=torch.float64)
        A = A @ A.T / n + torch.diag(torch.linspace(0.05, 4.0, n, dtype=torch.float64))
        c = torch.randn(n, generator=g, dtype=torch.float64)
        parts.append((A, c))
    init_shared = torch.randn(n_shared, generator=g, dtype=torch.float64) * 0.3
    init_local = [torch.randn(nl, generator=g, dtype=torch.float64) * 0.3 for nl in n_local]

    def f_r(r, shared, local):
        A, c = parts[r]
        x = torch.cat([shared.reshape(-1), local.reshape(-1)])
        return 0.5 * x @ (A @ x) - c @ x + 0.1 * torch.sum(torch.log1p(x * x))
    return f_r, init_shared, init_local


def _reference_run(steps, max_iter, history):
    f_r, s0, l0 = _objective()
    shared = torch.nn.Parameter(s0.clone().reshape(3, 4))
    local = [torch.nn.Parameter(t.clone()) for t in l0]
    opt = torch.optim.LBFGS([shared] + local, max_iter=max_iter, history_size=history)
    losses = []

    def closure():
        opt.zero_grad()
        loss = sum(f_r(r, shared, local[r]) for r in range(len(local)))
        loss.backward()
        losses.append(float(loss.detach()))
        return loss
    for _ in range(steps):
        opt.step(closure)
    return losses, shared.detach().clone(), [t.detach().clone() for t in local]


# ----------------------------------------------------------------------------- single process, torch backend
@pytest.mark.parametrize("steps,max_iter,history", [(1, 20, 100), (2, 6, 3), (3, 1, 100)])
def test_fused_lbfgs_host_logic_matches_torch_on_cpu_backend(steps, max_iter, history):
    from optim import FusedLBFGS

    class CpuLBFGS(TorchPasses, FusedLBFGS):
        pass
    f_r, s0, l0 = _objective()
    shared = torch.nn.Parameter(s0.clone().reshape(3, 4))
    local = [torch.nn.Parameter(t.clone()) for t in l0]
    opt = CpuLBFGS([shared] + local, max_iter=max_iter, history_size=history)
    losses = []

    def closure():
        opt.zero_grad()
        loss = sum(f_r(r, shared, local[r]) for r in range(len(local)))
        loss.backward()
        losses.append(float(loss.detach()))
        return loss
    for _ in range(steps):
        opt.step(closure)
    ref_losses, ref_shared, ref_local = _reference_run(steps, max_iter, history)
    assert len(losses) == len(ref_losses)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-9)
    np.testing.assert_allclose(shared.detach().numpy(), ref_shared.numpy(), rtol=1e-7, atol=1e-10)
    for a, b in zip(local, ref_local):
        np.testing.assert_allclose(a.detach().numpy(), b.numpy(), rtol=1e-7, atol=1e-10)


# ----------------------------------------------------------------------------- world_size 2 over gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, steps, max_iter, history, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from parallel import ShardedLBFGS, shard_sessions

        class CpuSharded(TorchPasses, ShardedLBFGS):
            pass
        assert shard_sessions(["s1", "s0", "s2"]) == (["s0", "s2"] if rank == 0 else ["s1"])
        f_r, s0, l0 = _objective()
        shared = torch.nn.Parameter(s0.clone().reshape(3, 4))          # replicated (the role V plays)
        local = torch.nn.Parameter(l0[rank].clone())                   # this rank's block (the role of U_s, b_s)
        opt = CpuSharded([shared], [local], max_iter=max_iter, history_size=history)
        losses = []

        def closure():
            opt.zero_grad()
            loss = f_r(rank, shared, local)
            loss.backward()
            buf = torch.cat([shared.grad.reshape(-1), loss.detach().reshape(1)])
            dist.all_reduce(buf)                                       # what parallel.joint_loss_and_grad does
            shared.grad.copy_(buf[:-1].view_as(shared.grad))
            losses.append(float(buf[-1]))
            return buf[-1]
        for _ in range(steps):
            opt.step(closure)
        torch.save({"losses": losses, "shared": shared.detach(), "local": local.detach()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("steps,max_iter,history", [(1, 20, 100), (2, 5, 3)])
def test_sharded_lbfgs_world2_equals_single_process(tmp_path, steps, max_iter, history):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, steps, max_iter, history, str(tmp_path)), nprocs=2, join=True)
    ref_losses, ref_shared, ref_local = _reference_run(steps, max_iter, history)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(2)]
    for i in range(2):
        np.testing.assert_allclose(r[i]["losses"], ref_losses, rtol=1e-9)
        np.testing.assert_allclose(r[i]["shared"].numpy(), ref_shared.numpy(), rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(r[i]["local"].numpy(), ref_local[i].numpy(), rtol=1e-7, atol=1e-10)
    assert torch.equal(r[0]["shared"], r[1]["shared"])                 # the replica never drifts


def test_shard_sessions_balanced_and_deterministic():
    from parallel import shard_sessions
    eids = [f"e{i}" for i in range(7)]
    cost = {e: c for e, c in zip(eids, [9, 1, 8, 2, 7, 3, 6])}
    parts = [shard_sessions(eids, r, 3, cost) for r in range(3)]
    assert sorted(sum(parts, [])) == sorted(eids)
    loads = [sum(cost[e] for e in p) for p in parts]
    assert max(loads) - min(loads) <= 3
    assert parts == [shard_sessions(list(reversed(eids)), r, 3, cost) for r in range(3)]
    assert [shard_sessions(eids, r, 2) for r in range(2)] == [eids[0::2], eids[1::2]]
