"""One steady-state step of a workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off ... python tools/profile_step.py --workload rrr|linear
(launch list with gpu__time_duration.sum, or --set full on one kernel).  Not a benchmark."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-spike_b200"))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rrr")
ap.add_argument("--trials", type=int, default=400)
ap.add_argument("--planes", type=int, default=1)
ap.add_argument("--evals", type=int, default=0, help="rrr: profile only this many closure evaluations (0 = whole fit)")
args = ap.parse_args()
dev = torch.device("cuda:0")
cudart = torch.cuda.cudart()

if args.workload == "rrr":
    from bench import rrr_inputs, sorted_idx_42
    from model.rrr import RRRGD, pack_session_from_frames, train_model
    from torch import optim
    ftr, ctr, fte, cte = rrr_inputs(args.trials, 80, 110 * 166, 144, 0, pinned=False)
    td = {"s": pack_session_from_frames(ftr, ctr, fte, cte, sorted_idx_42(), 3, planes=args.planes, device=dev)}
    model = RRRGD(td, 3, l2=100.0, planes=args.planes); model.to(dev)
    init = {k: v.detach().clone() for k, v in model.model.items()}

    def fit():
        with torch.no_grad():
            for k, v in init.items():
                model.model[k].copy_(v)
        train_model(model, td, optim.LBFGS(model.model.parameters()), "tmp", save=False)

    fit(); fit()
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    if args.evals:
        for _ in range(args.evals):
            model.loss_and_grad(td, 0)
    else:
        fit()
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()
else:
    from tests.helpers import make_linear_model
    B, D, N = 16, 120 * 128 * 128, 144
    model, opt, sched = make_linear_model(D, N, dev, total_steps=5000)
    fr = torch.randint(0, 9, (B, D), dtype=torch.uint8, device=dev)
    tg = torch.poisson(torch.full((B, 100, N), 0.3, device=dev))
    for _ in range(4):
        model.fused_train_step(fr, tg, opt)
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    for _ in range(2):
        model.fused_train_step(fr, tg, opt)
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()
print("profiled", args.workload)
