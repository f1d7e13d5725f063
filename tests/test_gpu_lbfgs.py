"""-m gpu: FusedLBFGS (device vector passes + host coefficient-space recursion) against torch.optim.LBFGS
driven by the same deterministic closure -- the optimiser src/model/rrr.py:177,199 uses."""
import numpy as np
import pytest
import torch

from tests.helpers import small_rrr_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vs(cuda):
    import vsb200
    return vsb200


def _problem(dev, seed=0):
    """Ill-conditioned smooth non-quadratic objective over three parameters of different shapes."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(7, 50, 3), (7, 1, 20), (3, 20)]
    init = [torch.randn(s, generator=g, dtype=torch.float64) * 0.3 for s in shapes]
    n = sum(int(np.prod(s)) for s in shapes)
    A = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (A @ A.T / n + torch.diag(torch.linspace(0.05, 5.0, n, dtype=torch.float64))).to(dev)
    c = torch.randn(n, generator=g, dtype=torch.float64).to(dev)

    def make():
        ps = [torch.nn.Parameter(t.clone().to(dev)) for t in init]

        def f():
            x = torch.cat([p.reshape(-1) for p in ps])
            return 0.5 * x @ (A @ x) - c @ x + 0.1 * torch.sum(torch.log1p(x * x))
        return ps, f
    return make


@pytest.mark.parametrize("history,steps,max_iter", [(100, 1, 20), (3, 1, 20), (5, 3, 7), (100, 2, 1)])
def test_fused_lbfgs_matches_torch(vs, cuda, history, steps, max_iter):
    from optim import FusedLBFGS
    make = _problem(cuda)
    traj = {}
    for name, cls in (("torch", torch.optim.LBFGS), ("fused", FusedLBFGS)):
        ps, f = make()
        opt = cls(ps, history_size=history, max_iter=max_iter)
        losses = []

        def closure():
            opt.zero_grad()
            loss = f()
            loss.backward()
            losses.append(float(loss))
            return loss
        for _ in range(steps):
            opt.step(closure)
        st = opt.state[ps[0]]
        traj[name] = (losses, [p.detach().cpu().numpy() for p in ps], st["func_evals"], st["n_iter"])
    lt, pt, et, it = traj["torch"]
    lf, pf, ef, itf = traj["fused"]
    assert (et, it) == (ef, itf) and len(lt) == len(lf)
    np.testing.assert_allclose(lf, lt, rtol=1e-9)
    for a, b in zip(pf, pt):
        np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-9)


def test_fused_lbfgs_converged_start_and_tolerances(vs, cuda):
    """opt_cond on the first evaluation returns immediately; tolerance_change stops like torch."""
    from optim import FusedLBFGS
    for cls in (torch.optim.LBFGS, FusedLBFGS):
        p = torch.nn.Parameter(torch.zeros(10, dtype=torch.float64, device=cuda))
        opt = cls([p])
        n = [0]

        def closure():
            opt.zero_grad()
            loss = (p * p).sum()
            loss.backward()
            n[0] += 1
            return loss
        opt.step(closure)
        assert n[0] == 1 and opt.state[p]["n_iter"] == 0
    # a quadratic is solved in a couple of iterations, then |loss - prev_loss| < 1e-9 ends the loop
    res = {}
    for name, cls in (("t", torch.optim.LBFGS), ("f", FusedLBFGS)):
        p = torch.nn.Parameter(torch.full((16,), 0.01, dtype=torch.float64, device=cuda))
        opt = cls([p], max_iter=50)
        n = [0]

        def closure():
            opt.zero_grad()
            loss = ((p - 0.02) ** 2).sum()
            loss.backward()
            n[0] += 1
            return loss
        opt.step(closure)
        res[name] = (n[0], opt.state[p]["n_iter"], p.detach().cpu().numpy())
    assert res["t"][:2] == res["f"][:2]
    np.testing.assert_allclose(res["f"][2], res["t"][2], rtol=1e-9)


def test_rrr_fit_same_with_both_optimisers(vs, cuda):
    """One RRR fit (rrr.py:164-202) driven by torch.optim.LBFGS and by FusedLBFGS through the same device closure."""
    from model.rrr import RRRGD, train_model
    from optim import FusedLBFGS
    td = small_rrr_problem(seed=7, K=30, Kt=10, F=150, N=12)
    out = {}
    for name, cls in (("torch", torch.optim.LBFGS), ("fused", FusedLBFGS)):
        m = RRRGD(td, 3, l2=100.0, planes=1)
        m.to(cuda)
        opt = cls(m.model.parameters())
        _, res = train_model(m, td, opt, "tmp", save=False)
        out[name] = (float(res["mse_val_mean"]), m.n_closure_evals, m.model["V"].detach().cpu().numpy(),
                     m.model["e1_U"].detach().cpu().numpy())
    assert out["torch"][1] == out["fused"][1] == 20
    assert out["fused"][0] == pytest.approx(out["torch"][0], rel=1e-8)
    np.testing.assert_allclose(out["fused"][2], out["torch"][2], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out["fused"][3], out["torch"][3], rtol=1e-6, atol=1e-9)
    # the parameters handed out by the model are still the ParameterDict entries the reference exposes
    assert set(m.model.keys()) == {"e1_U", "e1_b", "V"} and m.model["e1_U"].shape == (12, 150, 3)


def test_float32_history_tracks_float64_history(vs, cuda):
    """history_dtype=float32 halves the L-BFGS traffic; the optimisation path stays the same to ~1e-6."""
    from optim import FusedLBFGS
    make = _problem(cuda, seed=3)
    res = {}
    for name, hd in (("f64", torch.float64), ("f32", torch.float32)):
        ps, f = make()
        opt = FusedLBFGS(ps, history_dtype=hd)
        losses = []

        def closure():
            opt.zero_grad()
            loss = f()
            loss.backward()
            losses.append(float(loss.detach()))
            return loss
        opt.step(closure)
        res[name] = (losses, torch.cat([p.detach().reshape(-1) for p in ps]).cpu().numpy(), opt.state[ps[0]]["func_evals"])
    assert res["f32"][2] == res["f64"][2]
    np.testing.assert_allclose(res["f32"][0], res["f64"][0], rtol=1e-5)
    np.testing.assert_allclose(res["f32"][1], res["f64"][1], rtol=1e-4, atol=1e-6)


def test_rrr_fit_float32_history(vs, cuda):
    from model.rrr import RRRGD, train_model
    td = small_rrr_problem(seed=8, K=30, Kt=10, F=150, N=12)
    out = {}
    for hd in (torch.float64, torch.float32):
        m = RRRGD(td, 3, l2=100.0, planes=1)
        m.to(cuda)
        _, res = train_model(m, td, m.make_optimizer(history_dtype=hd), "tmp", save=False)
        out[hd] = float(res["mse_val_mean"])
    assert out[torch.float32] == pytest.approx(out[torch.float64], rel=1e-5)


# ----------------------------------------------------------------------------- device-driven mode (no host sync per iteration)
@pytest.mark.parametrize("history,steps,max_iter,hd", [(100, 1, 20, torch.float64), (3, 1, 20, torch.float64), (5, 3, 7, torch.float64),
                                                      (100, 2, 1, torch.float64), (100, 1, 20, torch.float32)])
def test_device_driven_lbfgs_matches_host_driven(vs, cuda, history, steps, max_iter, hd):
    """The one-thread update kernel takes the same decisions as the host logic: identical trajectories (to rounding),
    identical logical counters, across memory windows, several step() calls and both history dtypes."""
    from optim import FusedLBFGS
    make = _problem(cuda, seed=5)
    res = {}
    for mode in (False, True):
        ps, f = make()
        opt = FusedLBFGS(ps, history_size=history, max_iter=max_iter, history_dtype=hd, device_driven=mode)
        losses = []

        def closure():
            opt.zero_grad()
            loss = f()
            loss.backward()
            losses.append(loss.detach().clone())
            return loss
        for _ in range(steps):
            opt.step(closure)
        st = opt.state[ps[0]]
        res[mode] = ([float(l) for l in losses], torch.cat([p.detach().reshape(-1) for p in ps]).cpu().numpy(), st["func_evals"], st["n_iter"])
    assert res[True][2:] == res[False][2:]
    assert len(res[True][0]) == len(res[False][0])
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-9 if hd == torch.float64 else 1e-6)
    np.testing.assert_allclose(res[True][1], res[False][1], rtol=1e-7 if hd == torch.float64 else 1e-4, atol=1e-9)


def test_device_driven_lbfgs_terminates_like_torch(vs, cuda):
    """Early termination is decided on the device: the remaining launches are no-ops (parameters untouched) and the
    logical counters equal torch's, although the closure itself is still called max_iter times."""
    from optim import FusedLBFGS
    # converged at the first evaluation
    p = torch.nn.Parameter(torch.zeros(10, dtype=torch.float64, device=cuda))
    opt = FusedLBFGS([p], device_driven=True)
    calls = [0]

    def closure():
        opt.zero_grad(); loss = (p * p).sum(); loss.backward(); calls[0] += 1
        return loss
    opt.step(closure)
    assert opt.state[p]["func_evals"] == 1 and opt.state[p]["n_iter"] == 0 and opt.state[p]["dev_done"] == 1 and calls[0] == 20
    assert float(p.abs().max()) == 0.0
    # a quadratic: solved in a few iterations, then the loss tolerance stops it -- same point, same counters as torch
    out = {}
    for name, make in (("torch", lambda q: torch.optim.LBFGS([q], max_iter=50)), ("dev", lambda q: FusedLBFGS([q], max_iter=50, device_driven=True))):
        q = torch.nn.Parameter(torch.full((16,), 0.01, dtype=torch.float64, device=cuda))
        o = make(q)

        def cl():
            o.zero_grad(); loss = ((q - 0.02) ** 2).sum(); loss.backward()
            return loss
        o.step(cl)
        out[name] = (o.state[q]["func_evals"], o.state[q]["n_iter"], q.detach().cpu().numpy())
    assert out["dev"][:2] == out["torch"][:2]
    np.testing.assert_allclose(out["dev"][2], out["torch"][2], rtol=1e-9)


def test_rrr_fit_device_driven(vs, cuda):
    from model.rrr import RRRGD, train_model
    from optim import FusedLBFGS
    td = small_rrr_problem(seed=7, K=30, Kt=10, F=150, N=12)
    out = {}
    for mode in (False, True):
        m = RRRGD(td, 3, l2=100.0, planes=3); m.to(cuda)
        opt = FusedLBFGS(m.model.parameters(), device_driven=mode)
        _, res = train_model(m, td, opt, "tmp", save=False)
        out[mode] = (float(res["mse_val_mean"]), m.n_closure_evals, m.model["e1_U"].detach().cpu().numpy())
    assert out[True][1] == out[False][1] == 20
    assert out[True][0] == pytest.approx(out[False][0], rel=1e-8)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-6, atol=1e-9)


# ----------------------------------------------------------------------------- compact history (one stored vector per evaluation)
@pytest.mark.parametrize("steps,max_iter", [(1, 20), (3, 7), (2, 1), (1, 50)])
def test_compact_history_lbfgs_matches_torch(vs, cuda, steps, max_iter):
    """FusedLBFGS(compact=True): the basis {g_0, y_0, y_1, ...} + its Gram matrix instead of the (s_i, y_i) pairs.
    Same trajectory as torch.optim.LBFGS (to rounding), same logical counters, over several step() calls."""
    from optim import FusedLBFGS
    make = _problem(cuda, seed=5)
    res = {}
    for name in ("torch", "compact"):
        ps, f = make()
        opt = torch.optim.LBFGS(ps, max_iter=max_iter) if name == "torch" else FusedLBFGS(ps, max_iter=max_iter, device_driven=True, compact=True)
        losses = []

        def closure():
            opt.zero_grad()
            loss = f()
            loss.backward()
            losses.append(loss.detach().clone())
            return loss
        for _ in range(steps):
            opt.step(closure)
        st = opt.state[ps[0]]
        res[name] = ([float(l) for l in losses], torch.cat([p.detach().reshape(-1) for p in ps]).cpu().numpy(), st["func_evals"], st["n_iter"])
    assert res["compact"][2:] == res["torch"][2:]
    k = len(res["torch"][0])                       # the device-driven loop keeps calling the closure after termination (no-ops)
    np.testing.assert_allclose(res["compact"][0][:k], res["torch"][0], rtol=1e-8)
    np.testing.assert_allclose(res["compact"][1], res["torch"][1], rtol=1e-6, atol=1e-9)


def test_rrr_fit_compact_history(vs, cuda):
    """One RRR fit (rrr.py:164-202) with the compact-history optimiser against the (s, y)-pair device-driven one."""
    from model.rrr import RRRGD, train_model
    from optim import FusedLBFGS
    td = small_rrr_problem(seed=7, K=30, Kt=10, F=150, N=12)
    out = {}
    for compact in (False, True):
        m = RRRGD(td, 3, l2=100.0, planes=3); m.to(cuda)
        opt = FusedLBFGS(m.model.parameters(), device_driven=True, compact=compact)
        _, res = train_model(m, td, opt, "tmp", save=False)
        out[compact] = (float(res["mse_val_mean"]), m.n_closure_evals, m.model["e1_U"].detach().cpu().numpy())
    assert out[True][1] == out[False][1] == 20
    assert out[True][0] == pytest.approx(out[False][0], rel=1e-8)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-6, atol=1e-9)
    with pytest.raises(vs.VsError):
        FusedLBFGS(m.model.parameters(), device_driven=False, compact=True)
