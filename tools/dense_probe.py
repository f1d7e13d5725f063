"""Probe of the dense RRR backward: one closure evaluation with the dense route on/off on a small synthetic session."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "video-spike_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from helpers import small_rrr_problem
from model.rrr import RRRGD

K, F, N = (int(a) for a in sys.argv[1:4])
td = small_rrr_problem(seed=K + N, K=K, Kt=5, F=F, N=N)
m = RRRGD(td, 3, l2=100.0, planes=1, engine=2); m.to("cuda")
out = {}
for mode in ("0", "1"):
    os.environ["VS_RRR_DENSE"] = mode
    loss = float(m.loss_and_grad(td, 0))
    torch.cuda.synchronize()
    out[mode] = m.model["e1_U"].grad.cpu().numpy().copy()
    print("mode", mode, "loss", loss, flush=True)
s = np.abs(out["0"]).max()
print(f"K={K} F={F} N={N}: max|dense-fact|/max|g| = {np.abs(out['1'] - out['0']).max() / s:.3e}", flush=True)
