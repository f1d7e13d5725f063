"""Phase breakdown of the RRR end-to-end fit (bench.py's e2e path) with a synchronize after each phase."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch
import bench
import vsb200 as vs
from model.rrr import RRRGD, pack_session_from_frames, train_model
from optim import FusedLBFGS

K, Kt, F, N = 400, 80, 110 * 166, 144
MODE = os.environ.get("MODE", "exact")            # exact | dense | classic (1 plane)
PL = 1 if MODE == "classic" else None
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, 0, pinned=True)
sidx = bench.sorted_idx_42()
dev = torch.device("cuda")
for it in range(3):
    T = {}
    def tick(name, t0):
        torch.cuda.synchronize(); T[name] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); a = ftr.to(dev, non_blocking=True); b = fte.to(dev, non_blocking=True); tick("h2d_frames", t0)
    del a, b
    t0 = time.perf_counter(); entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=PL, mode=MODE); tick("pack(h2d+R0)", t0)
    td = {"s": entry}
    t0 = time.perf_counter(); model = RRRGD(td, 3, l2=100.0, planes=PL); tick("RRRGD.__init__", t0)
    t0 = time.perf_counter(); model.to(dev); tick("to(device)", t0)
    t0 = time.perf_counter(); opt = model.make_optimizer(); _, res = train_model(model, td, opt, "tmp", save=False); v = float(res["mse_val_mean"]); tick("fit+val", t0)
    ms = torch.cuda.memory_stats()
    print(it, {k: round(v, 2) for k, v in T.items()}, "sum", round(sum(list(T.values())[1:]), 1),
          "| cudaMalloc calls so far", ms.get("num_device_alloc"), "frees", ms.get("num_device_free"), "reserved GB", round(ms["reserved_bytes.all.current"] / 2**30, 2))
    del model, opt, td, entry, res
# same thing through the one-call public API, timed as bench.py's e2e does
from model.rrr import train_model_from_frames
import gc
gc_events = []
gc.callbacks.append(lambda phase, info: gc_events.append((phase, info.get('generation'), time.perf_counter())))
for it in range(12):
    if it == 6: gc.collect(); gc.disable(); print('--- gc disabled')
    n_ev = len(gc_events)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m, res, _ = train_model_from_frames(ftr, ctr, fte, cte, sidx, l2=100.0, n_comp=3, planes=PL, mode=MODE)
    v = float(res["mse_val_mean"]); dt = (time.perf_counter() - t0) * 1e3
    ms = torch.cuda.memory_stats()
    print("api", it, round(dt, 2), "ms | gc events", [(p, g) for p, g, _ in gc_events[n_ev:]], "| cudaMalloc", ms.get("num_device_alloc"), "frees", ms.get("num_device_free"), "reserved GB", round(ms["reserved_bytes.all.current"] / 2**30, 2))
    del m, res, _
