"""Along the trajectory of ONE full-size fit driven by mode A, evaluate the closure of mode B at the same parameters and
print the per-evaluation differences (loss, gradients) -- finds precision holes that the start/end probes miss.
Also checks bit-reproducibility of mode B.   A=exact B=dense SEED=1 python tools/mode_diff_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch
import bench
from model.rrr import RRRGD, pack_session_from_frames

dev = torch.device("cuda")
K, Kt, F, N = 400, 80, 110 * 166, 144
A, B = os.environ.get("A", "exact"), os.environ.get("B", "dense")
seed = int(os.environ.get("SEED", 1))
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, seed, pinned=False)
sidx = bench.sorted_idx_42()
ea = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, device=dev, mode=A)
eb = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, device=dev, mode=B)
ma = RRRGD({"s": ea}, 3, l2=100.0); ma.to(dev)
mb = RRRGD({"s": eb}, 3, l2=100.0); mb.to(dev)
opt = ma.make_optimizer()
it = [0]

# float64 truth at the same parameters (the reference's formulation: einsum + autograd)
from scipy.ndimage import gaussian_filter1d
sid = torch.as_tensor(np.asarray(sidx), device=dev)
X = ftr.to(dev).reshape(K, 120, -1).double()
mean = X.mean(0); std = X.std(0, unbiased=False).clamp_min(1e-8)
Xtr = torch.cat([(X - mean) / std, torch.ones(K, 120, 1, dtype=torch.float64, device=dev)], 2)[:, sid]
del X
ys = gaussian_filter1d(ctr.numpy().astype(np.float64), 2, axis=1)
ytr = torch.from_numpy((ys - ys.mean(0)) / np.clip(ys.std(0), 1e-8, None)).to(dev)


def truth():
    U = ma.model["s_U"].detach().clone().requires_grad_(True); V = ma.model["V"].detach().clone().requires_grad_(True)
    b = ma.model["s_b"].detach().clone().requires_grad_(True)
    beta = torch.cat([U @ V, b], 1)
    l = ((torch.einsum("ktc,nct->ktn", Xtr, beta) - ytr) ** 2).sum() + 100.0 * (beta ** 2).sum()
    l.backward()
    return float(l), {"U": U.grad, "b": b.grad, "V": V.grad}

def closure():
    opt.zero_grad()
    la = ma.loss_and_grad({"s": ea}, 0)
    with torch.no_grad():
        for k in ma.model:
            mb.model[k].copy_(ma.model[k])
    lb = mb.loss_and_grad({"s": eb}, 0)
    gb = {k: mb.model[k].grad.clone() for k in mb.model}
    lb2 = mb.loss_and_grad({"s": eb}, 0)
    same = float(lb) == float(lb2) and all(torch.equal(gb[k], mb.model[k].grad) for k in gb)
    lt, gt = truth()
    ea_ = {k.split("_")[-1]: float((ma.model[k].grad - gt[k.split("_")[-1]]).norm() / gt[k.split("_")[-1]].norm()) for k in ma.model}
    eb_ = {k.split("_")[-1]: float((gb[k] - gt[k.split("_")[-1]]).norm() / gt[k.split("_")[-1]].norm()) for k in ma.model}
    print(f"        vs float64: |gV| {float(gt['V'].norm()):.3e} |gU| {float(gt['U'].norm()):.3e} | {A}: loss {abs(float(la)-lt)/lt:.1e} U {ea_['U']:.1e} V {ea_['V']:.1e} "
          f"| {B}: loss {abs(float(lb)-lt)/lt:.1e} U {eb_['U']:.1e} V {eb_['V']:.1e}", flush=True)
    d = {k.split("_")[-1]: float((mb.model[k].grad - ma.model[k].grad).norm() / ma.model[k].grad.norm()) for k in ma.model}
    dm = {k.split("_")[-1]: float((mb.model[k].grad - ma.model[k].grad).abs().max() / ma.model[k].grad.abs().max()) for k in ma.model}
    print(f"eval {it[0]:2d} loss {float(la):.6e} | {B} vs {A}: loss rel {abs(float(lb)-float(la))/abs(float(la)):.2e} grad rel-L2 U {d['U']:.2e} b {d['b']:.2e} V {d['V']:.2e} "
          f"| max-abs/max U {dm['U']:.2e} V {dm['V']:.2e} | |U|max {float(ma.model['s_U'].abs().max()):.3f} |V|max {float(ma.model['V'].abs().max()):.3f} | {B} reproducible {same}", flush=True)
    it[0] += 1
    return la

opt.step(closure)
