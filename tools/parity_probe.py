"""Full-size RRR fit (BASELINE configs[1]) in every operand mode against an independent float64 dense implementation
(bench.fp64_dense_reference: torch einsum + autograd + torch.optim.LBFGS on the GPU in fp64, the reference's own formulation).
MODES=exact,exact32,p3,p2,p1f16,p1  selects what runs (default: all)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch
import bench
from model.rrr import RRRGD, pack_session_from_frames

dev = torch.device("cuda")
K, Kt, F, N = int(os.environ.get("K", 400)), 80, 110 * 166, 144
seed = int(os.environ.get("SEED", 0))
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, seed, pinned=False, signal=os.environ.get("SIGNAL", "0") == "1")
sidx = bench.sorted_idx_42()
t0 = time.time()
ref = bench.fp64_dense_reference(ftr, ctr, fte, cte, sidx, dev)
ref_val = ref["val_sse"]
print(f"fp64 dense reference (seed {seed}): val SSE {ref_val:.6f}  evals {ref['evals']}  ({time.time()-t0:.1f} s)  first/last train loss "
      f"{ref['first_loss']:.6e} {ref['last_loss']:.6e}", flush=True)
ALL = {"exact": ("exact", None, None, torch.float64), "exact32": ("exact", None, None, torch.float32),
       "dense": ("dense", None, None, torch.float64), "dense32": ("dense", None, None, torch.float32),
       "p3": ("classic", 3, "bf16", torch.float64), "p2": ("classic", 2, "bf16", torch.float64),
       "p2f16": ("classic", 2, "f16", torch.float64),
       "p1f16": ("classic", 1, "f16", torch.float32), "p1": ("classic", 1, "bf16", torch.float32)}
names = (os.environ.get("MODES") or ",".join(ALL)).split(",")
for nm in names:
    mode, planes, operand, hd = ALL[nm]
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=planes, device=dev, operand=operand, mode=mode)
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0, planes=planes, operand=operand); m.to(dev)
    # per-evaluation error at the perturbed start of the reference
    with torch.no_grad():
        m.model["s_U"].copy_(ref["start"]["U"]); m.model["s_b"].copy_(ref["start"]["b"]); m.model["V"].copy_(ref["start"]["V"])
    l1 = float(m.loss_and_grad(td, 0)); lref, gref = ref["probe"]
    ge = {k: float((m.model[n_].grad - gref[k]).norm() / gref[k].norm()) for k, n_ in (("U", "s_U"), ("b", "s_b"), ("V", "V"))}
    # the same at the END point of the reference fit (small gradient)
    with torch.no_grad():
        m.model["s_U"].copy_(ref["end"]["U"]); m.model["s_b"].copy_(ref["end"]["b"]); m.model["V"].copy_(ref["end"]["V"])
    l2_ = float(m.loss_and_grad(td, 0)); lref2, gref2 = ref["probe_end"]
    ge2 = {k: float((m.model[n_].grad - gref2[k]).norm() / gref2[k].norm()) for k, n_ in (("U", "s_U"), ("b", "s_b"), ("V", "V"))}
    gn2 = {k: float(gref2[k].norm()) for k in gref2}
    m2 = RRRGD(td, 3, l2=100.0, planes=planes, operand=operand); m2.to(dev)
    losses = []
    o = m2.make_optimizer(history_dtype=hd)
    def cl():
        o.zero_grad(); l = m2.loss_and_grad(td, 0); losses.append(l); return l
    torch.cuda.synchronize(); t1 = time.perf_counter()
    o.step(cl)
    val = float(torch.sum(m2.compute_MSE_RRRGD(td, 1)["s"]))
    fit_ms = (time.perf_counter() - t1) * 1e3
    print(f"{nm:8s} hist={str(hd)[6:]:8s} val SSE {val:.4f} rel diff {abs(val-ref_val)/ref_val:.3e} | per-eval loss rel {abs(l1-lref)/abs(lref):.2e} "
          f"grad rel-L2 U {ge['U']:.2e} b {ge['b']:.2e} V {ge['V']:.2e} | END point: loss rel {abs(l2_-lref2)/abs(lref2):.2e} grad rel-L2 "
          f"U {ge2['U']:.2e} b {ge2['b']:.2e} V {ge2['V']:.2e} (|g| U {gn2['U']:.2e} b {gn2['b']:.2e} V {gn2['V']:.2e}) | fit {fit_ms:.1f} ms", flush=True)
    del m, m2, td, entry, o
    torch.cuda.empty_cache()
