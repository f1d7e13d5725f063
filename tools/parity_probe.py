"""Full-size RRR fit (BASELINE configs[1]) in every operand mode against an independent float64 dense implementation
(torch einsum + autograd + torch.optim.LBFGS on the GPU in fp64, i.e. the reference's own formulation)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch
from scipy.ndimage import gaussian_filter1d
import bench
from model.rrr import RRRGD, pack_session_from_frames, train_model
from optim import FusedLBFGS

dev = torch.device("cuda")
K, Kt, F, N = int(os.environ.get("K", 400)), 80, 110 * 166, 144
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, 0, pinned=False, signal=os.environ.get("SIGNAL", "1") == "1")
sidx = bench.sorted_idx_42()

# ---- float64 dense reference on the GPU (formulation of src/model/rrr.py:79-155, preprocessing of train_rrr.py:108-171)
def prep(fr, mean=None, std=None):
    X = fr.to(dev).double()
    if mean is None:
        mean = X.mean(0); std = X.std(0, unbiased=False).clamp_min(1e-8)
    X = (X - mean) / std
    X = torch.cat([X, torch.ones(X.shape[0], X.shape[1], 1, dtype=torch.float64, device=dev)], 2)
    return X[:, torch.as_tensor(sidx, device=dev)], mean, std
Xtr, mX, sX = prep(ftr); Xte, _, _ = prep(fte, mX, sX)
ytr = gaussian_filter1d(ctr.numpy().astype(np.float64), 2, axis=1); yte = gaussian_filter1d(cte.numpy().astype(np.float64), 2, axis=1)
my, sy = ytr.mean(0), np.clip(ytr.std(0), 1e-8, None)
ytr = torch.from_numpy((ytr - my) / sy).to(dev); yte = torch.from_numpy((yte - my) / sy).to(dev)
np.random.seed(0)
U0 = np.random.normal(size=(N, F, 3)) / np.sqrt(300); V0 = np.random.normal(size=(3, 100)) / np.sqrt(300)
U = torch.nn.Parameter(torch.from_numpy(U0).to(dev)); V = torch.nn.Parameter(torch.from_numpy(V0).to(dev))
b = torch.nn.Parameter(ytr.mean(0).T.unsqueeze(1).contiguous())
opt = torch.optim.LBFGS([U, b, V])
trace = []
def closure():
    opt.zero_grad()
    beta = torch.cat([U @ V, b], 1)                                   # (N, C, T)
    pred = torch.einsum("ktc,nct->ktn", Xtr, beta)
    loss = ((pred - ytr) ** 2).sum() + 100.0 * (beta ** 2).sum()
    loss.backward(); trace.append(float(loss.detach())); return loss
t0 = time.time(); opt.step(closure); torch.cuda.synchronize()
with torch.no_grad():
    beta = torch.cat([U @ V, b], 1)
    ref_val = float(((torch.einsum("ktc,nct->ktn", Xte, beta) - yte) ** 2).sum())
print(f"fp64 dense reference: val SSE {ref_val:.6f}  evals {len(trace)}  ({time.time()-t0:.1f} s)  first/last train loss {trace[0]:.6e} {trace[-1]:.6e}")
del Xtr, Xte, beta
torch.cuda.empty_cache()
MODES = ((3, "bf16", torch.float64), (2, "bf16", torch.float64), (1, "f16", torch.float64), (1, "f16", torch.float32),
         (1, "bf16", torch.float64), (1, "bf16", torch.float32))
if os.environ.get("FAST_ONLY"):
    MODES = ((1, "f16", torch.float32), (1, "bf16", torch.float32))
for planes, operand, hd in MODES:
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=planes, device=dev, operand=operand)
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0, planes=planes, operand=operand); m.to(dev)
    losses = []
    o = m.make_optimizer(history_dtype=hd)
    def cl():
        o.zero_grad(); l = m.loss_and_grad(td, 0); losses.append(float(l)); return l
    torch.cuda.synchronize(); t1 = time.perf_counter()
    o.step(cl)
    val = float(torch.sum(m.compute_MSE_RRRGD(td, 1)["s"]))
    fit_ms = (time.perf_counter() - t1) * 1e3
    dev_tr = max(abs(a - r) / abs(r) for a, r in zip(losses, trace))
    print(f"planes={planes} {operand:4s} hist={str(hd)[6:]:8s} val SSE {val:.4f} rel diff {abs(val-ref_val)/ref_val:.3e} | first-eval loss rel diff {abs(losses[0]-trace[0])/trace[0]:.3e} max over evals {dev_tr:.3e} | fit {fit_ms:.1f} ms  VS_RRR_RUN={os.environ.get('VS_RRR_RUN')}")
    del m, td, entry, o
    torch.cuda.empty_cache()
