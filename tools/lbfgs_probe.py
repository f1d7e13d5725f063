"""Times the two L-BFGS passes (float64 vs float32 history) and one whole fit per history dtype."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import ctypes as C
import numpy as np, torch
import vsb200 as vs
dev = torch.device("cuda")
n = 144 * 18260 * 3 + 144 * 100 + 300
for hd in (torch.float64, torch.float32):
    for m in (4, 10, 19):
        npad = (n + 3) // 4 * 4
        hist = torch.randn((2 * m + 2, npad), device=dev, dtype=torch.float64).to(hd)
        g = torch.randn(n, device=dev, dtype=torch.float64); gp = torch.randn_like(g); x = torch.randn_like(g)
        out = torch.zeros(8 + 6 * m + 8, device=dev, dtype=torch.float64)
        ws = torch.empty(int(vs.lib.vs_lbfgs_workspace(n, m)), dtype=torch.uint8, device=dev)
        ss = (C.c_int32 * m)(*range(0, 2 * m, 2)); ys = (C.c_int32 * m)(*range(1, 2 * m, 2))
        coef = (C.c_double * (2 * m + 1))(*([0.1] * (2 * m + 1)))
        f32 = int(hd == torch.float32)
        def dots():
            vs.check(vs.lib.vs_lbfgs_dots(n, vs.ptr(g), vs.ptr(gp), vs.ptr(hist[2 * m]), vs.ptr(hist[2 * m + 1]), vs.ptr(hist), npad, f32, ss, ys, m,
                                          vs.ptr(out), vs.ptr(ws), ws.numel(), vs.stream()))
        def direction():
            vs.check(vs.lib.vs_lbfgs_direction(n, vs.ptr(g), vs.ptr(hist), npad, f32, ss, ys, m, coef, 1.0, vs.ptr(x), vs.ptr(hist[2 * m]), vs.ptr(out[-2:-1]), vs.stream()))
        for name, fn in (("dots", dots), ("direction", direction)):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            esz = 4 if f32 else 8
            byts = n * ((16 + esz + esz + 2 * m * esz) if name == "dots" else (8 + 2 * m * esz + esz + 16))
            print(f"{hd} m={m:2d} {name:9s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.1f} GB/s")
import bench
from model.rrr import RRRGD, pack_session_from_frames, train_model
K, Kt, F, N = 400, 80, 110 * 166, 144
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, 0, pinned=True)
entry = pack_session_from_frames(ftr, ctr, fte, cte, bench.sorted_idx_42(), 3, planes=1)
td = {"s": entry}
model = RRRGD(td, 3, l2=100.0, planes=1); model.to(dev)
init = {k: v.detach().clone() for k, v in model.model.items()}
for hd in (torch.float64, torch.float32, torch.float64, torch.float32):
    ts = []
    for it in range(4):
        with torch.no_grad():
            for k, v in init.items(): model.model[k].copy_(v)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        opt = model.make_optimizer(history_dtype=hd)
        _, res = train_model(model, td, opt, "tmp", save=False)
        v = float(res["mse_val_mean"]); ts.append((time.perf_counter() - t0) * 1e3)
    print(hd, "fit ms", [round(t, 2) for t in ts], "val", v)
# where does a fit spend host time?  closure-only loop (20 evals, no optimiser)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): l = model.loss_and_grad(td, 0)
float(l); print("20 closure evals back to back: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
for _ in range(20): l = model.loss_and_grad(td, 0); float(l)
print("20 closure evals with a sync each: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
