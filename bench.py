#!/usr/bin/env python
"""bench.py -- training video frames/s of the video->spike hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload both|rrr|linear] [--impl reference]

ONE JSON line.  The top level is the RRR workload (config.workload); the Linear workload is nested under "linear".
  rrr     BASELINE configs[1] at N = 1: RRR rank-3, ONE session, K=400 train trials of 120 frames x (110 x 166) whisker
          ROI, N=144 neurons, l2=100.  A "step" is one full fit (src/model/rrr.py:164-202: one LBFGS.step = 20 closure
          evaluations + the validation pass).  frames/s = K*120 / fit time.  Default operand mode "exact" (the mode whose
          whole fit is within 1e-3 of the float64 reference; --mode classic --planes 1 is the faster non-parity mode).
          N > 1 GPUs (BASELINE configs[2]): ONE joint model over N sessions, one per GPU, with a shared V
          (rrr.py:37-49): [dV, loss] are all-reduced per closure evaluation and the L-BFGS scalars all-gathered per
          iteration (NCCL, stream-ordered: the optimiser stays device-driven).  --independent: N separate fits instead
          (what train_rrr.py:179-187 runs; no data-path collective).  --strong: ONE session with its trials sharded over
          the N GPUs (gradient all-reduce per evaluation; "scaling": "strong").
  linear  BASELINE configs[0]/[3]: `Linear` MLP train step (src/trainer/base.py:147-154), B=16, D=120*128*128, N=144:
          frames/s = B*120 / step time.  N > 1: row-parallel first layer (each rank owns 1/N of the pixels of W0 and of
          every frame; one 16 KB all-reduce per step), strong scaling.

`value`  : device-timed (CUDA events), inputs resident in HBM.
`e2e`    : same metric through the public Python API from PINNED HOST buffers, H2D/D2H inside the timed region: the MEAN
           of a fixed number of calls (median and every sample listed beside it).
`roofline`: dominant kernel, event-bracketed inside the timed region (vs_profile_*), vs MEASURED_PEAKS.json.
`parity` : the timed mode against an independent float64 dense fit (torch einsum + autograd + torch.optim.LBFGS on the GPU).
`cpu_baseline`: the oracle (CPU port of the reference) timed on the host cores on a bounded sample: whole fits on all
           trials and features for a sample of the NEURONS, carried to the full neuron count by an affine model calibrated
           in the run (`extrapolated`, `factor_on_value`, `calibration` are in the line).
--impl reference: only that CPU arm, printed in the same JSON shape (rank 0 alone under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "video-spike_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC, UNIT = "training video frames/sec", "frames/s"
FRAMES_PER_TRIAL = 120
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


def ncu_traffic(key, field="bytes_per_launch_avg"):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/), or None."""
    try:
        ent = None
        for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                with open(path) as fh:
                    ent = json.load(fh).get(key)
            if ent:
                break
        return int(ent[field]) if ent and field in ent else None            # bytes per launch
    except Exception:
        return None


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.t0 = [], None, gpu_index, None

    def start(self):
        """Launch the sampling process (before the warm-up: nvidia-smi needs ~0.1 s to deliver its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout=1.5):
        """Block until the sampling process has delivered its first line (so that it is running when the timed region starts)."""
        t = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def begin(self):
        """Start of the timed region: only samples taken from here on are reported."""
        self.t0 = time.perf_counter()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        t1 = time.perf_counter()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        t0 = self.t0 if self.t0 is not None else 0.0
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.02]
        in_region = len(rows)
        if not rows:     # timed region shorter than one sampling period: the samples of the warm-up steps just before it (same load)
            rows = [r for (ts, r) in self.rows if ts >= t0 - 0.5]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        smax = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(sm), "samples_in_timed_region": in_region}


# ----------------------------------------------------------------------------- distributed plumbing
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# ----------------------------------------------------------------------------- synthetic inputs
def rrr_inputs(K, Kt, F, N, seed, pinned, signal=False):
    """Synthetic session of the loader's shape: dense uniform 0..255 uint8 frames (K,120,1,h,w
    flattened to F) and Poisson(0.4) spike counts (K,100,N).  signal=True plants a rank-3 dependence of the rate on the
    frames instead (used by tools/parity_probe.py: on it the reference's un-line-searched L-BFGS is chaotic even in
    float64 -- profiles/r01_parity_probe_signal_data.txt).  Host tensors, optionally pinned."""
    g = torch.Generator().manual_seed(seed)
    mk = (lambda *s, **k: torch.empty(*s, **k).pin_memory()) if pinned else torch.empty
    ftr = mk((K, FRAMES_PER_TRIAL, F), dtype=torch.uint8); ftr.random_(0, 256, generator=g)
    fte = mk((Kt, FRAMES_PER_TRIAL, F), dtype=torch.uint8); fte.random_(0, 256, generator=g)
    ctr = mk((K, 100, N), dtype=torch.float32)
    cte = mk((Kt, 100, N), dtype=torch.float32)
    if not signal:
        ctr.copy_(torch.poisson(torch.full((K, 100, N), 0.4), generator=g))
        cte.copy_(torch.poisson(torch.full((Kt, 100, N), 0.4), generator=g))
        return ftr, ctr, fte, cte
    Wt = torch.randn((F, 3), generator=g) / float(np.sqrt(F))
    Vt = torch.randn((3, 100), generator=g)
    A = torch.randn((3, N), generator=g)
    for fr, cnt in ((ftr, ctr), (fte, cte)):
        for k0 in range(0, fr.shape[0], 32):                     # chunked: the float copy of the frames is 4x their size
            x = (fr[k0:k0 + 32, :100].float() - 127.5) / 74.0     # frames 0..99 drive bins 0..99
            lat = torch.einsum("ktf,fj->ktj", x, Wt) * Vt.T.unsqueeze(0)
            rate = torch.exp(torch.clamp(0.3 * torch.einsum("ktj,jn->ktn", lat, A), -3.0, 2.0) - 1.0)
            cnt[k0:k0 + 32].copy_(torch.poisson(rate, generator=g))
    return ftr, ctr, fte, cte


def sorted_idx_42():
    st = np.random.get_state()
    np.random.seed(42)                      # utils.set_seed(42) then the first numpy draw (train_rrr.py:41,48-49)
    idx = np.sort(np.random.choice(119, 100, replace=False))
    np.random.set_state(st)
    return idx


# ----------------------------------------------------------------------------- CPU arm (oracle)
def _cpu_rrr_problem(K, Kt, F, N_s, seed=0):
    """The reference's float64 train_data (src/train_rrr.py:108-171) for ALL K trials and all F features but only N_s
    neurons.  Frame selection first (the statistics of a frame depend on that frame alone), in place, to bound host memory."""
    from scipy.ndimage import gaussian_filter1d
    sidx = sorted_idx_42()
    ftr, ctr, fte, cte = rrr_inputs(K, Kt, F, N_s, seed, pinned=False)
    Xs = []
    mean = std = None
    for fr in (ftr, fte):
        x = np.empty((fr.shape[0], len(sidx), F + 1), dtype=np.float64)
        x[:, :, :F] = fr.numpy()[:, sidx]
        if mean is None:
            mean = x[:, :, :F].mean(0); std = np.clip(x[:, :, :F].std(0), 1e-8, None)
        x[:, :, :F] -= mean; x[:, :, :F] /= std
        x[:, :, F] = 1.0
        Xs.append(x)
    ys = [gaussian_filter1d(c.numpy().astype(np.float64), 2, axis=1) for c in (ctr, cte)]
    my, sy = ys[0].mean(0), np.clip(ys[0].std(0), 1e-8, None)
    return {"s": {"X": Xs, "y": [(y - my) / sy for y in ys]}}


def _cpu_rrr_fit(td):
    """One whole fit of the reference on the host cores (src/model/rrr.py:164-202): the oracle's autograd transcription of
    the closure (beta built twice, X/y converted on every call, einsum + backward) driven by the oracle's restatement of
    torch.optim.LBFGS.step, then the validation pass on split 1.  Returns (seconds, closure evaluations, val SSE)."""
    from oracle import rrr_oracle as ro
    params = ro.rrr_init(td, 3)
    order = list(params.keys())
    t0 = time.perf_counter()
    n_eval = [0]

    def closure(x):
        n_eval[0] += 1
        loss, grads = ro.loss_and_grad_autograd(ro._unflatten(x, params, order), td, 100.0)
        return loss, ro._flatten(grads, order)

    x, _ = ro.lbfgs_step(closure, ro._flatten(params, order))
    p = ro._unflatten(x, params, order)
    beta = ro.compute_beta(p["s_U"], p["V"], p["s_b"])
    val = float(np.sum((ro.predict(beta, td["s"]["X"][1]) - td["s"]["y"][1]) ** 2))
    return time.perf_counter() - t0, n_eval[0], val


def cpu_rrr_sample(F, N, n_fits, n_warm, budget_s, K, Kt):
    """Host-core baseline on a BOUNDED sample of the workload: ALL K trials and all F features, but only N_s of the N neurons.
    The terms of the reference's fit scale with the neuron count (beta = (N, C, T) built twice per closure evaluation, the
    einsum K*T*C*N and its backward, the L-BFGS vectors N*C*r) up to a per-evaluation constant (einsum's handling of X):
    two calibration evaluations (4 and 12 neurons) give the affine model  eval_s = a + b * neurons,  N_s is chosen from it so
    that (n_warm + n_fits) whole fits take about budget_s seconds, and the measured fit time is carried to N neurons by
    (a + b N) / (a + b N_s).  (Sampling TRIALS instead would leave the beta term, ~2 s per evaluation at N = 144, unsampled:
    a 4-trial fit takes as long as a 40-trial one.)"""
    from oracle import rrr_oracle as ro
    torch.set_num_threads(os.cpu_count() or 1)

    full = _cpu_rrr_problem(K, Kt, F, N)                          # X does not depend on the neuron count: built once

    def sub(n):
        return {"s": {"X": full["s"]["X"], "y": [np.ascontiguousarray(y[:, :, :n]) for y in full["s"]["y"]]}}

    def eval_s(n):
        td = sub(n)
        p = ro.rrr_init(td, 3)
        ro.loss_and_grad_autograd(p, td, 100.0)                  # warm-up (allocator, threads)
        t0 = time.perf_counter()
        ro.loss_and_grad_autograd(p, td, 100.0)
        return time.perf_counter() - t0

    n_lo, n_hi = min(N, 4), min(N, 12)
    t_lo, t_hi = eval_s(n_lo), eval_s(n_hi)
    b = max((t_hi - t_lo) / max(n_hi - n_lo, 1), 1e-9)           # seconds per evaluation per neuron
    a = max(t_lo - b * n_lo, 0.0)                                # seconds per evaluation independent of the neuron count
    per_fit = budget_s / (n_warm + n_fits)
    N_s = int((per_fit / 22.0 - a) / b)
    N_s = max(1, min(N, N_s))
    td = sub(N_s)
    for _ in range(n_warm):
        _cpu_rrr_fit(td)
    secs, evals = [], 0
    for _ in range(n_fits):
        dt, evals, _ = _cpu_rrr_fit(td)
        secs.append(dt)
    fit_s = float(np.mean(secs))
    factor = (a + b * N_s) / (a + b * N)                          # measured fit time -> full neuron count (affine model)
    return {"value": K * FRAMES_PER_TRIAL / fit_s * factor, "fit_s": fit_s, "fit_s_each": [round(x, 2) for x in secs], "neurons": N_s,
            "evals": evals, "fits_timed": n_fits, "fits_warmup": n_warm, "factor": factor,
            "calibration": {f"eval_s_at_{n_lo}_neurons": t_lo, f"eval_s_at_{n_hi}_neurons": t_hi, "a_s_per_eval": a, "b_s_per_eval_per_neuron": b},
            "fit_s_full_size_estimate": fit_s / factor}


def cpu_linear_sample(B, D, N, steps=2):
    from oracle import linear_oracle as lo
    torch.set_num_threads(os.cpu_count() or 1)
    tr = lo.Trainer(lo.init_params(D, N, seed=42), total_steps=5000)
    frames, ap = lo.synth_batch(B, (D,), N, seed=0, dist="sparse")
    tr.step(frames, ap)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(frames, ap)
    dt = (time.perf_counter() - t0) / steps
    return B * FRAMES_PER_TRIAL / dt, dt


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on the host cores (the oracle port: the reference is Python and needs
    packages this image lacks).  Every step is a WHOLE fit / train step, on a bounded sample."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload != "linear":
        r = cpu_rrr_sample(args.features, args.neurons, max(1, args.steps), max(0, args.warmup), args.ref_budget, args.trials, args.trials_test)
        sample = (f"oracle whole fits (CPU port of src/model/rrr.py:164-202: autograd closure, torch fp64, + L-BFGS vector work + validation "
                  f"pass) on ALL {args.trials} train / {args.trials_test} val trials x 120 frames at full C={args.features + 1}, for {r['neurons']} of the "
                  f"{args.neurons} neurons: {r['fits_warmup']} warm-up + {r['fits_timed']} timed fits of {r['evals']} closure evaluations, "
                  f"{r['fit_s']:.2f} s per fit; carried to {args.neurons} neurons with the affine model eval_s = a + b*neurons calibrated in "
                  f"this run ({r['calibration']}): value = measured frames/s x {r['factor']:.4f} (full-size fit estimate "
                  f"{r['fit_s_full_size_estimate']:.0f} s)")
        cfg = rrr_config(args, world)
        cfg.update({"neurons": r["neurons"], "neurons_full": args.neurons, "sessions": 1, "parallelism": f"host cores x{cores}",
                    "operand_mode": "float64 (reference)", "operand_planes": None, "operand_format": "float64",
                    "lbfgs": "1 step, max_iter 20 (oracle restatement of torch.optim.LBFGS)",
                    "extrapolated": r["neurons"] < args.neurons,
                    "extrapolation": {"sampled": "neurons", "run": r["neurons"], "full": args.neurons, "factor_on_value": r["factor"],
                                      "model": "eval_s = a + b * neurons, two calibration evaluations in this run", "calibration": r["calibration"]}})
        v, ms = r["value"], r["fit_s"] * 1e3
    else:
        v, dt = cpu_linear_sample(args.batch, args.input_dim, args.neurons, steps=max(1, min(args.steps, 3)))
        sample = f"oracle Trainer.step (CPU port of src/trainer/base.py:147-154, torch fp32), B={args.batch}, D={args.input_dim}, N={args.neurons}"
        cfg = linear_config(args, world)
        ms = dt * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.workload == "linear" else "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- configs
def rrr_mode_of(args):
    if args.mode == "classic" or (args.mode is None and args.planes is not None):
        return "classic"
    return args.mode or os.environ.get("VS_RRR_MODE") or "exact"


def rrr_config(args, world):
    mode = rrr_mode_of(args)
    exact = mode in ("exact", "dense")
    strong = world > 1 and getattr(args, "strong", False)
    joint = world > 1 and not args.independent and not strong
    planes = 2 if exact else (args.planes or 1)
    if mode == "dense":
        fmt = ("dense: both contractions per time bin on the exact integer frames (half): forward x hi+lo coefficient planes generated on chip, "
               "backward and dV pass x hi+lo residual planes; fp32 accumulate in TMEM, float64 epilogues")
    elif exact:
        fmt = ("exact: forward operand = z-score as hi+lo IEEE-half planes (3 plane products), backward operand = exact integer frames "
               "(half) x hi+lo residual planes, fp32 accumulate in TMEM, float64 epilogues")
    else:
        fmt = (os.environ.get("VS_RRR_OPERAND") or "bf16") + f" x {planes} plane(s) (16-bit tensor-core operands, fp32 accumulate in TMEM)"
    hist = "float64" if planes > 1 else "float32"
    return {"workload": "rrr_joint_fit, shared V across sessions (BASELINE configs[2])" if joint else
                        ("rrr_single_session_fit (BASELINE configs[1]), trials sharded over the GPUs" if strong else "rrr_single_session_fit (BASELINE configs[1])"),
            "trials_train": args.trials, "trials_test": args.trials_test,
            "frames_per_trial": FRAMES_PER_TRIAL, "frame_shape": "1x110x166", "features": args.features, "time_bins": 100,
            "neurons": args.neurons, "rank": 3, "l2": 100, "operand_mode": mode, "operand_planes": planes,
            "operand_format": fmt,
            "lbfgs": f"1 step, max_iter 20 (20 closure evals), history {hist}, "
                     + ("device-driven" if os.environ.get("VS_LBFGS_DEVICE", "1") != "0" else "host-driven")
                     + (", compact history (one vector per evaluation)" if (hist == "float64" and os.environ.get("VS_LBFGS_DEVICE", "1") != "0"
                                                                             and os.environ.get("VS_LBFGS_COMPACT", "1") != "0") else "")
                     + (", inner products sharded over the ranks" if joint else ""),
            "sessions": 1 if strong else world,
            "parallelism": (f"ONE session, trials sharded over {world} GPUs ({args.trials // world} train trials each), parameters and L-BFGS state replicated: "
                            f"per closure evaluation ONE NCCL all-reduce of the flat gradient ({(args.neurons * (args.features + 1) * 3 + 300 + args.neurons * 100) * 8 / 1e6:.0f} MB float64) and one of the loss"
                            if strong else f"joint model over {world} sessions, one per GPU: shared V; per closure evaluation ONE NCCL all-reduce of [dV, loss], "
                            f"per L-BFGS iteration ONE NCCL all-gather of the {8 + 6 * 100 + 1} optimiser scalars (stream-ordered, no host sync)"
                            if joint else (f"independent sessions, one per GPU x{world} (no collective)" if world > 1 else "single GPU")),
            "l2_cache": "operands (>= 1.46 GB per pass) exceed the 126 MB L2; no flush needed"}


def linear_config(args, world):
    return {"workload": "linear_mlp_train_step (BASELINE configs[0] on B200)", "batch": args.batch, "input_dim": args.input_dim,
            "neurons": args.neurons, "optimizer": "AdamW+OneCycleLR",
            "parallelism": f"first layer row-parallel over {world} GPUs (pixel shards), one 16 KB all-reduce per step" if world > 1 else "single GPU",
            "l2_cache": "weights+Adam state (6 GB) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- float64 dense reference on the GPU
def fp64_dense_reference(ftr, ctr, fte, cte, sidx, dev, perturb_seed=1):
    """The reference's own formulation (src/model/rrr.py:79-155,164-202; preprocessing of src/train_rrr.py:108-171) in
    float64 torch on the GPU: einsum + autograd + torch.optim.LBFGS.  Independent of every kernel of this repo.
    Returns {"val_sse", "evals", "probe": (loss, {name: grad}) at a perturbed start, "start": {name: tensor}}."""
    from scipy.ndimage import gaussian_filter1d
    sid = torch.as_tensor(np.asarray(sidx), device=dev)

    def prep(fr, mean=None, std=None):
        X = fr.to(dev).reshape(fr.shape[0], fr.shape[1], -1).double()
        if mean is None:
            mean = X.mean(0); std = X.std(0, unbiased=False).clamp_min(1e-8)
        X = (X - mean) / std
        X = torch.cat([X, torch.ones(X.shape[0], X.shape[1], 1, dtype=torch.float64, device=dev)], 2)
        return X[:, sid], mean, std

    Xtr, mX, sX = prep(ftr)
    Xte, _, _ = prep(fte, mX, sX)
    ytr = gaussian_filter1d(ctr.numpy().astype(np.float64), 2, axis=1)
    yte = gaussian_filter1d(cte.numpy().astype(np.float64), 2, axis=1)
    my, sy = ytr.mean(0), np.clip(ytr.std(0), 1e-8, None)
    ytr = torch.from_numpy((ytr - my) / sy).to(dev); yte = torch.from_numpy((yte - my) / sy).to(dev)
    N, F = ytr.shape[2], Xtr.shape[2] - 1
    st = np.random.get_state()
    np.random.seed(0)
    U0 = np.random.normal(size=(N, F, 3)) / np.sqrt(300); V0 = np.random.normal(size=(3, 100)) / np.sqrt(300)
    np.random.set_state(st)
    U = torch.nn.Parameter(torch.from_numpy(U0).to(dev)); V = torch.nn.Parameter(torch.from_numpy(V0).to(dev))
    b = torch.nn.Parameter(ytr.mean(0).T.unsqueeze(1).contiguous())

    def loss_fn():
        beta = torch.cat([U @ V, b], 1)                                   # (N, C, T)
        pred = torch.einsum("ktc,nct->ktn", Xtr, beta)
        return ((pred - ytr) ** 2).sum() + 100.0 * (beta ** 2).sum()

    # (1) one evaluation at a perturbed start (at the exact init the gradient of b is ~0 by construction)
    init = {"U": U.detach().clone(), "b": b.detach().clone(), "V": V.detach().clone()}
    gp = torch.Generator().manual_seed(perturb_seed)
    start = {k: v + (0.02 * torch.randn(v.shape, generator=gp, dtype=torch.float64)).to(dev) for k, v in init.items()}
    with torch.no_grad():
        U.copy_(start["U"]); b.copy_(start["b"]); V.copy_(start["V"])
    l = loss_fn(); l.backward()
    probe = (float(l), {"U": U.grad.clone(), "b": b.grad.clone(), "V": V.grad.clone()})
    with torch.no_grad():
        U.copy_(init["U"]); b.copy_(init["b"]); V.copy_(init["V"])
    # (2) the whole fit
    opt = torch.optim.LBFGS([U, b, V])
    trace = []

    def closure():
        opt.zero_grad()
        l = loss_fn(); l.backward(); trace.append(float(l.detach())); return l

    opt.step(closure)
    with torch.no_grad():
        beta = torch.cat([U @ V, b], 1)
        pred = torch.einsum("ktc,nct->ktn", Xte, beta)
        val = float(((pred - yte) ** 2).sum())
        pred_fr = (pred.cpu().numpy() * sy + my)              # src/model/rrr.py:136-142 (predict_y_fr)
    del beta, pred
    # (3) one more evaluation at the END of the fit, where the gradient is small (dV is then a small difference of large sums)
    end = {"U": U.detach().clone(), "b": b.detach().clone(), "V": V.detach().clone()}
    opt.zero_grad()
    l = loss_fn(); l.backward()
    probe_end = (float(l.detach()), {"U": U.grad.clone(), "b": b.grad.clone(), "V": V.grad.clone()})
    del Xtr, Xte
    torch.cuda.empty_cache()
    return {"val_sse": val, "evals": len(trace), "probe": probe, "start": start, "first_loss": trace[0], "last_loss": trace[-1],
            "pred_test_fr": pred_fr, "end": end, "probe_end": probe_end}


# ----------------------------------------------------------------------------- RRR workload
def run_rrr(args, rank, world, local):
    import vsb200 as vs
    from model.rrr import RRRGD, pack_session_from_frames, train_model, train_model_from_frames
    vs.require_b200()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    K, Kt, F, N = args.trials, args.trials_test, args.features, args.neurons
    mode = rrr_mode_of(args)
    planes = None if mode == "exact" else (args.planes or 1)
    sidx = sorted_idx_42()
    strong = world > 1 and args.strong            # ONE session, its trials sharded over the ranks (every rank generates the same session)
    ftr, ctr, fte, cte = rrr_inputs(K, Kt, F, N, seed=0 if strong else rank, pinned=True)
    if strong:
        cut = lambda n: (n * rank // world, n * (rank + 1) // world)
        (ka, kb), (ta, tb) = cut(K), cut(Kt)
        ftr_l, ctr_l, fte_l, cte_l = ftr[ka:kb], ctr[ka:kb], fte[ta:tb], cte[ta:tb]
    # bytes that cross PCIe per fit: the 100 selected frames of every trial (vs_h2d_select_frames) + the spike counts
    h2d = (ftr.numel() + fte.numel()) // ftr.shape[1] * len(sidx) + 4 * (ctr.numel() + cte.numel())
    if strong:
        from parallel import pack_trial_shard, build_trial_sharded_model, fit_trial_sharded
        h2d = (ftr_l.numel() + fte_l.numel()) // ftr.shape[1] * len(sidx) + 4 * (ctr_l.numel() + cte_l.numel())

    # ---- resident-input measurement: operands packed once, fit repeated
    if strong:
        entry = pack_trial_shard(ftr_l, ctr_l, fte_l, cte_l, sidx, 3, planes=planes or 1, device=dev, mode=mode)
    else:
        entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=planes, device=dev, mode=mode)
    # N > 1 (BASELINE configs[2]): ONE model over all ranks' sessions with a shared V (src/model/rrr.py:37-49); rank r holds
    # session r's U, b and operands; [dV, loss] all-reduced per evaluation, the L-BFGS scalars all-gathered per iteration
    joint = world > 1 and not args.independent and not strong
    eid = f"s{rank:02d}" if joint else "s"
    plan = [(f"s{r:02d}", N, F + 1, 100) for r in range(world)] if joint else None
    td = {eid: entry}
    if strong:
        model = build_trial_sharded_model(entry, 100.0, 3, eid=eid)
    else:
        model = RRRGD(td, 3, l2=100.0, planes=planes, init_plan=plan)
        model.to(dev)
    model_fmt, model_planes = model.fmt, model.planes
    init = {k: v.detach().clone() for k, v in model.model.items()}
    if joint:
        from parallel import train_joint_model

    def one_fit():
        with torch.no_grad():
            for k, v in init.items():
                model.model[k].copy_(v)
        if strong:
            _, res = fit_trial_sharded(model, entry, eid=eid)
        elif joint:
            _, res = train_joint_model(model, td)
        else:
            opt = model.make_optimizer()                   # what train_model_main builds (rrr.py:199 semantics)
            _, res = train_model(model, td, opt, "tmp", save=False)
        return res["mse_val_mean"]

    sampler = ClockSampler(local); sampler.start(); sampler.wait_first()
    for _ in range(args.warmup):
        one_fit()
    torch.cuda.synchronize(); barrier(world)
    vs.lib.vs_launch_count_reset(); vs.lib.vs_profile_enable(1)
    evals0 = model.n_closure_evals
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    e0.record()
    for _ in range(args.steps):
        mse = one_fit()
    e1.record(); torch.cuda.synchronize(); barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / args.steps
    clocks = sampler.stop()
    launches = int(vs.lib.vs_launch_count())
    n_gemm, gemm_ms, gmin, gmax = vs.profile_read(0)          # forward GEMMs (train closures + the evaluation split)
    n_bwd, bwd_ms, _, _ = vs.profile_read(2)                  # dense backward kernel (0 launches when VS_RRR_DENSE=0)
    n_fwd, fwd_ms, _, _ = vs.profile_read(3)                  # dense forward kernel (dense mode): train closures + the evaluation split
    n_dv, dv_ms, _, _ = vs.profile_read(4)                    # dV pass of the dense backward kernel (dense mode)
    vs.lib.vs_profile_enable(0)
    evals = (model.n_closure_evals - evals0) / args.steps
    n_sessions = 1 if strong else world
    value = n_sessions * K * FRAMES_PER_TRIAL / (ms * 1e-3)
    K_full, Kt_full = K, Kt
    if strong:                                      # the kernels below ran on this rank's shard of the trials
        K, Kt = kb - ka, tb - ta

    # roofline of the dominant kernels: algorithmic FLOPs (2*K*T*C*N per contraction, i.e. the dense formulation)
    # or bytes, summed over the launches of the timed region / their summed event-bracketed duration.
    C = F + 1
    Np16 = (N + 15) // 16 * 16
    Kp = (K + 15) // 16 * 16
    pk, pk_kind = peaks()
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    key = f"rrr_K{K}_F{F}_N{N}_" + (mode if mode != "classic" else f"planes{model_planes}")
    tu = "bytes/launch (ncu dram read+write of an EARLIER run of this command, profiles/r02_ncu_traffic.json; not measured in this run)"
    r_planes = 2 if mode in ("exact", "dense") else 1
    bwd_bytes = 2.0 * F * 100 * Kp + 2.0 * r_planes * Np16 * 100 * Kp + 4.0 * F * 3 * Np16      # operand + R planes + G out
    blocks = []
    if mode == "dense":
        fl = args.steps * (evals * 2.0 * K * 100 * F * N + 2.0 * Kt * 100 * F * N)
        ach = fl / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0
        blocks.append({"bound": "tensor", "kernel": "vs::tc::rrr_fwd_dense_pair_kernel (yhat_t = Xc_t beta'_t^T per time bin: exact integer A by TMA, hi+lo "
                                                    "coefficient tiles generated on chip, tcgen05 cta_group::2 UMMA 256xN, 3 bins per CTA pair)",
                       "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                       "traffic": ncu_traffic(key, "fwd_bytes_per_launch"), "traffic_unit": tu, "peak_source": f"{pk_kind} bf16_tflops_sustained",
                       "launches": n_fwd, "avg_launch_ms": fwd_ms / max(n_fwd, 1), "share_of_step": fwd_ms / (ms * args.steps),
                       "executed_tflops": 2.0 * ach * (256.0 * ((K + 255) // 256) / K),
                       "hbm_view": {"algorithmic_bytes_per_launch": 2.0 * K * 100 * F, "achieved_gbs": 2.0 * K * 100 * F * args.steps * evals / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0,
                                    "peak_gbs": pk["hbm_gbs"]},
                       "note": "achieved counts the ALGORITHMIC flops 2*K*T*C*N once; the kernel executes two plane products (hi, lo) on 256-row tiles (executed_tflops)"})
        if n_dv > 0:
            bw = bwd_bytes * n_dv / (dv_ms * 1e-3) / 1e9
            blocks.append({"bound": "hbm", "kernel": "vs::tc::rrr_bwd_dense_pair_kernel<dV> (dV: the D_t = Xc_t^T R_t tiles contracted with the U slab held in registers/TMEM)",
                           "achieved": bw, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bw / pk["hbm_gbs"], "traffic": ncu_traffic(key, "dv_bytes_per_launch"),
                           "launches": n_dv, "avg_launch_ms": dv_ms / n_dv, "share_of_step": dv_ms / (ms * args.steps)})
    else:
        plane_passes = 1 if model_planes == 1 else (3 if model_planes == 2 else 6)
        n_contr = 1 if n_bwd > 0 else 2                             # contractions per closure evaluation under tag 0
        algo_flops = args.steps * (evals * n_contr * (2.0 * K * 100 * C * N) + 2.0 * Kt * 100 * C * N)
        exec_flops = args.steps * (evals * n_contr * (2.0 * K * 100 * F * 3 * Np16) + 2.0 * Kt * 100 * F * 3 * Np16) * plane_passes
        ach = algo_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        blocks.append({"bound": "tensor", "kernel": "vs::tc::gemm_tn_pair_kernel (forward Z = X U; tcgen05 cta_group::2 kind::f16, UMMA 256xN over a CTA pair"
                                                    + (f"; {plane_passes} plane products per launch)" if plane_passes > 1 else ")"),
                       "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                       "traffic": ncu_traffic(key), "traffic_unit": tu, "peak_source": f"{pk_kind} bf16_tflops_sustained",
                       "launches": n_gemm, "avg_launch_ms": gemm_ms / max(n_gemm, 1), "share_of_step": gemm_ms / (ms * args.steps),
                       "executed_tflops": exec_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
                       "executed_frac": (exec_flops / (gemm_ms * 1e-3) / 1e12 / peak) if (gemm_ms > 0 and peak) else None,
                       "algorithmic_bytes_per_launch": 2.0 * K * 100 * (F + 4) * (2 if model_planes >= 2 else 1)      # A planes once
                                                       + 2.0 * 3 * Np16 * (F + 4) * (2 if model_planes >= 2 else 1)          # B planes
                                                       + 4.0 * K * 100 * 3 * Np16 * (-(-(-(-F // 64)) // int(os.environ.get('VS_RRR_RUN_EXACT', '72') or 72)) if mode == 'exact' else 1),   # Z partial tiles (accumulation runs) out
                       "note": "achieved counts ALGORITHMIC flops of the dense formulation (2*K*T*C*N per contraction); the factorised forward executes "
                               f"r=3x that per plane product and {plane_passes} plane product(s) (executed_tflops, executed_frac: the kernel is "
                               "bound by the tensor pipe, 88.8 % active under ncu).  traffic exceeds the algorithmic bytes because the hi plane of X "
                               "is read by two of the three plane products (DESIGN.md section 8)"})
    if n_bwd > 0:
        bw = bwd_bytes * n_bwd / (bwd_ms * 1e-3) / 1e9
        blocks.append({"bound": "hbm", "kernel": "vs::tc::rrr_bwd_dense_pair_kernel (dU: D_t = X_t^T R_t per time bin in TMEM, rank-one updates in registers"
                                                 + ("; exact integer operand x hi+lo residual planes)" if r_planes == 2 else ")"),
                       "achieved": bw, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bw / pk["hbm_gbs"],
                       "traffic": ncu_traffic(key, "bwd_bytes_per_launch"), "launches": n_bwd, "avg_launch_ms": bwd_ms / n_bwd,
                       "share_of_step": bwd_ms / (ms * args.steps),
                       "algorithmic_tflops": 2.0 * K * 100 * C * N * n_bwd / (bwd_ms * 1e-3) / 1e12})
    # the top-level object is the kernel with the largest share of the step; the others ride along under "other_kernels"
    blocks.sort(key=lambda x: -x["share_of_step"])
    roof = blocks[0]
    roof["other_kernels"] = blocks[1:]

    # ---- parity of the timed configuration (outside every timed region) against an INDEPENDENT float64 dense fit on the
    # GPU (the reference's formulation in torch: einsum + autograd + torch.optim.LBFGS), same session
    parity = None
    if rank == 0 and not args.no_parity and strong:
        ref = fp64_dense_reference(ftr, ctr, fte, cte, sidx, dev)
        parity = {"reference": "float64 dense fit of the WHOLE session on one GPU (torch einsum + autograd + torch.optim.LBFGS; src/model/rrr.py:79-155,164-202)",
                  "fit_val_sse": float(mse), "fit_val_sse_fp64": ref["val_sse"], "fit_rel_diff": abs(float(mse) - ref["val_sse"]) / ref["val_sse"],
                  "tolerance": 1e-3, "within_tolerance": bool(abs(float(mse) - ref["val_sse"]) / ref["val_sse"] <= 1e-3), "fp64_evals": ref["evals"]}
        del ref
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_parity and not joint and not strong:
        ref = fp64_dense_reference(ftr, ctr, fte, cte, sidx, dev)
        with torch.no_grad():
            model.model["s_U"].copy_(ref["start"]["U"]); model.model["s_b"].copy_(ref["start"]["b"]); model.model["V"].copy_(ref["start"]["V"])
        l1 = float(model.loss_and_grad(td, 0))
        lref, gref = ref["probe"]
        names = {"U": "s_U", "b": "s_b", "V": "V"}
        gerr = {k: float((model.model[nm].grad - gref[k]).abs().max() / gref[k].abs().max()) for k, nm in names.items()}
        gerr2 = {k: float((model.model[nm].grad - gref[k]).norm() / gref[k].norm()) for k, nm in names.items()}
        parity = {"reference": "float64 dense fit on the GPU (torch einsum + autograd + torch.optim.LBFGS; src/model/rrr.py:79-155,164-202)",
                  "per_eval_loss_rel_diff": abs(l1 - lref) / abs(lref), "per_eval_grad_max_abs_diff_over_max_abs": gerr,
                  "per_eval_grad_rel_l2": gerr2,
                  "fit_val_sse": float(mse), "fit_val_sse_fp64": ref["val_sse"], "fit_rel_diff": abs(float(mse) - ref["val_sse"]) / ref["val_sse"],
                  "tolerance": 1e-3, "within_tolerance": bool(abs(float(mse) - ref["val_sse"]) / ref["val_sse"] <= 1e-3),
                  "fp64_evals": ref["evals"]}
        del ref, gref
        torch.cuda.empty_cache()

    # ---- end to end: pinned host uint8 frames -> R0 on device -> init -> fit -> validation loss on the host
    import contextlib

    def e2e_fit():
        with contextlib.redirect_stdout(sys.stderr):         # the reference's prints ("GPU is available", ...) stay off the JSON stream
            return _e2e_fit()

    def _e2e_fit():
        if strong:
            ent = pack_trial_shard(ftr_l, ctr_l, fte_l, cte_l, sidx, 3, planes=planes or 1, device=dev, mode=mode)
            m = build_trial_sharded_model(ent, 100.0, 3, eid=eid)
            _, res = fit_trial_sharded(m, ent, eid=eid)
        elif joint:
            ent = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=planes, device=dev, mode=mode)
            m = RRRGD({eid: ent}, 3, l2=100.0, planes=planes, init_plan=plan, device=dev)
            _, res = train_joint_model(m, {eid: ent})
        else:
            m, res, _ = train_model_from_frames(ftr, ctr, fte, cte, sidx, l2=100.0, n_comp=3, planes=planes, mode=mode)
        return float(res["mse_val_mean"])                    # device -> host read of the result

    del model, td, entry
    torch.cuda.empty_cache()
    import gc
    for _ in range(max(3, args.warmup)):            # warm-up calls: the caching allocator re-grows after the parity reference's empty_cache()
        e2e_fit()
    gc.collect(); gc.disable()                      # no collector pauses inside the timed region (re-enabled below)
    torch.cuda.synchronize(); barrier(world)
    n_e2e = max(10, args.steps)                     # FIXED number of calls: no adaptive stopping
    each = []
    for i in range(n_e2e):
        t1 = time.perf_counter()
        val = e2e_fit()
        each.append((time.perf_counter() - t1) * 1e3)
        gc.collect()                                # between fits, outside the per-fit timing: frees the previous fit's operands
    torch.cuda.synchronize(); barrier(world)
    gc.enable()
    mean_s = max_over_ranks(float(np.mean(each)) * 1e-3, world, dev)
    med_s = max_over_ranks(float(np.median(each)) * 1e-3, world, dev)
    K, Kt = K_full, Kt_full
    e2e = {"value": n_sessions * K * FRAMES_PER_TRIAL / mean_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
           "ms_per_step": mean_s * 1e3, "statistic": f"mean of {n_e2e} consecutive calls (max over ranks)",
           "median_ms_per_step": med_s * 1e3, "value_at_median": n_sessions * K * FRAMES_PER_TRIAL / med_s,
           "ms_each_rank0": [round(x, 2) for x in each],
           "path": ("parallel.pack_trial_shard(this rank's trials, pinned uint8 frames) -> build_trial_sharded_model -> fit_trial_sharded -> float(mse_val_mean)" if strong
                    else "pack_session_from_frames(pinned uint8 frames) -> RRRGD(init_plan: own U and the kept V drawn every call; the other ranks' sessions are jumped over with stream positions remembered from the warm-up calls) -> parallel.train_joint_model -> float(mse_val_mean)" if joint
                    else "model.rrr.train_model_from_frames(pinned uint8 frames) -> float(mse_val_mean)")}

    # ---- the reference's own entry point: train_model_main(train_data) with the float64 numpy arrays train_rrr.py builds
    # (5.8 GB + 1.2 GB of float64 over PCIe, packed into 3 residual planes on the device): the drop-in call, few samples
    if rank == 0 and world == 1 and args.dropin_e2e > 0:
        try:
            e2e["dropin_fp64"] = dropin_fp64_e2e(ftr, ctr, fte, cte, sidx, args.dropin_e2e, K)
        except Exception as exc:                     # never lose the line over the optional measurement
            e2e["dropin_fp64"] = {"error": repr(exc)[:200]}

    if rank != 0:
        return None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_rrr_sample(F, N, 1, 0, args.cpu_budget, K, Kt)
        cpu = {"value": r["value"], "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle whole fit (CPU port of src/model/rrr.py:164-202, torch fp64 autograd closure + L-BFGS + validation pass) on ALL "
                         f"{K} train / {Kt} val trials x 120 frames at full C={C}, for {r['neurons']} of the {N} neurons: 1 fit of {r['evals']} closure "
                         f"evaluations, {r['fit_s']:.1f} s; carried to {N} neurons with eval_s = a + b*neurons calibrated in this run: value = measured x {r['factor']:.4f}",
               "extrapolated": r["neurons"] < N, "neurons_run": r["neurons"], "factor_on_value": r["factor"], "calibration": r["calibration"],
               "fit_s_full_size_estimate": r["fit_s_full_size_estimate"]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f16" if model_fmt == 1 else "bf16",
            "data": "synthetic", "config": rrr_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "closure_evals_per_step": evals, "val_sse": float(mse), "val_sse_e2e": val, "parity": parity}
    return line


def dropin_fp64_e2e(ftr, ctr, fte, cte, sidx, n, K):
    """`model.rrr.train_model_main(train_data, l2, n_comp, fname, save=False)` with the reference's own inputs: float64
    numpy X (K, 100, C) / y (K, 100, N) as src/train_rrr.py:143-171 builds them (host preprocessing NOT timed)."""
    from scipy.ndimage import gaussian_filter1d
    from model.rrr import train_model_main
    Xs = [f.numpy().reshape(f.shape[0], f.shape[1], -1) for f in (ftr, fte)]
    mean = Xs[0].mean(0, dtype=np.float64); std = np.clip(Xs[0].std(0, dtype=np.float64), 1e-8, None)
    ys = [gaussian_filter1d(c.numpy().astype(np.float64), 2, axis=1) for c in (ctr, cte)]
    my, sy = ys[0].mean(0), np.clip(ys[0].std(0), 1e-8, None)
    Xo = []
    for x in Xs:
        z = np.empty((x.shape[0], len(sidx), x.shape[2] + 1), dtype=np.float64)
        z[:, :, :-1] = (x[:, sidx] - mean[sidx]) / std[sidx]
        z[:, :, -1] = 1.0
        Xo.append(z)
    td = {"s": {"X": Xo, "y": [(y - my) / sy for y in ys], "setup": {"mean_X_Tv": mean, "std_X_Tv": std, "mean_y_TN": my, "std_y_TN": sy}}}
    nbytes = sum(a.nbytes for a in Xo) + sum(a.nbytes for a in td["s"]["y"])
    import contextlib, io
    each, val = [], None
    for i in range(n + 1):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            _, res = train_model_main(td, 100.0, 3, "tmp", save=False)
        val = float(res["mse_val_mean"])
        if i:
            each.append((time.perf_counter() - t0) * 1e3)
    mean_ms = float(np.mean(each))
    return {"value": K * FRAMES_PER_TRIAL / (mean_ms * 1e-3), "unit": UNIT, "ms_per_step": mean_ms, "ms_each": [round(x, 1) for x in each],
            "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": 8, "val_sse": val,
            "path": "model.rrr.train_model_main(float64 numpy train_data) -> float(mse_val_mean)  [default operand planes for host arrays: "
                    + os.environ.get("VS_RRR_PLANES", "3") + "]"}


# ----------------------------------------------------------------------------- Linear workload
def make_linear_model(input_dim, n_neurons, device, total_steps=5000, seed=42):
    """Model + optimizer + OneCycleLR built the way src/train.py:36-57 builds them from the YAML configs."""
    from model.linear import Linear
    from optim import FusedAdamW
    from utils.config_utils import config_from_kwargs, update_config
    cfg = config_from_kwargs({"model": "include:" + os.path.join(PKG, "config", "model", "linear_video.yaml")})
    cfg = update_config(os.path.join(PKG, "config", "train", "linear_video.yaml"), cfg)
    cfg["model"]["encoder"]["input_dim"] = input_dim
    cfg["model"]["decoder"]["output_dim"] = 100 * n_neurons
    torch.manual_seed(seed)
    model = Linear(cfg.model).to(device)
    o = cfg.optimizer
    opt = FusedAdamW(model.parameters(), lr=o.lr, weight_decay=o.wd, eps=o.eps)
    sched = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, total_steps=total_steps, max_lr=o.lr, pct_start=o.warmup_pct,
                                                div_factor=o.div_factor)
    return model, opt, sched


def run_linear(args, rank, world, local, steps):
    import vsb200 as vs
    vs.require_b200()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, D, N = args.batch, args.input_dim, args.neurons
    model, opt, sched = make_linear_model(D, N, dev, total_steps=5000)
    lo_px, hi_px = 0, D
    if world > 1:
        # row-parallel first layer: rank g keeps W0[:, D_g] and receives only the pixels D_g of every frame;
        # ONE all-reduce of the (B, 256) pre-activation per step; the global batch and the optimisation are those of 1 GPU
        from optim import FusedAdamW
        lo_px, hi_px = model.shard_first_layer(rank, world)
        o = opt.defaults
        opt = FusedAdamW(model.parameters(), lr=o["lr"], weight_decay=o["weight_decay"], eps=o["eps"])
        sched = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, total_steps=5000, max_lr=5e-5, pct_start=0.15, div_factor=10)
        torch.cuda.empty_cache()
    Dl = hi_px - lo_px
    g = torch.Generator().manual_seed(0)                    # the SAME batches on every rank (each keeps its pixel slice)
    nbuf = 4
    frames_h = [torch.empty((B, Dl), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    ap_h = [torch.empty((B, 100, N), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    for f, a in zip(frames_h, ap_h):
        full = torch.empty((B, D), dtype=torch.uint8)
        full.random_(0, 9, generator=g)        # small values: the reference diverges on dense 0..255 frames
        f.copy_(full[:, lo_px:hi_px])
        a.copy_(torch.poisson(torch.full((B, 100, N), 0.3), generator=g))
    frames_d = [f.to(dev) for f in frames_h]
    ap_d = [a.to(dev) for a in ap_h]
    train_step = model.fused_train_step_rowpar if world > 1 else model.fused_train_step

    def step(i, fr, ap):
        loss = train_step(fr[i % nbuf], ap[i % nbuf], opt)
        sched.step()
        return loss

    import warnings
    warnings.filterwarnings("ignore", message="Detected call of `lr_scheduler.step")
    sampler = ClockSampler(local); sampler.start(); sampler.wait_first()
    for i in range(max(args.warmup, 3)):
        step(i, frames_d, ap_d)
    torch.cuda.synchronize(); barrier(world)
    vs.lib.vs_launch_count_reset(); vs.lib.vs_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    e0.record()
    for i in range(steps):
        loss = step(i, frames_d, ap_d)
    e1.record(); torch.cuda.synchronize(); barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / steps
    clocks = sampler.stop()
    launches = int(vs.lib.vs_launch_count())
    n_k, k_ms, _, _ = vs.profile_read(1)
    vs.lib.vs_profile_enable(0)
    value = B * FRAMES_PER_TRIAL / (ms * 1e-3)              # ONE global batch per step whatever the world size (strong scaling)
    P0 = Dl * 256
    algo_bytes = n_k * (24.0 * P0 + B * Dl + 4.0 * B * 256)    # p,m,v read+write, frames once, dH1 (this rank's slice)
    pk, pk_kind = peaks()
    ach = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "vs::dw_adamw_kernel (fused first-layer dW + AdamW)", "achieved": ach, "peak": pk["hbm_gbs"],
            "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic(f"linear_B{B}_D{D}_N{N}") if world == 1 else None,
            "traffic_unit": "bytes/launch (ncu dram read+write of an EARLIER run, profiles/r01_ncu_traffic.json; not measured in this run)",
            "peak_source": f"{pk_kind} hbm_gbs", "launches": n_k,
            "avg_launch_ms": k_ms / max(n_k, 1), "share_of_step": k_ms / (ms * steps),
            "step_roofline": {"algorithmic_bytes": 28.0 * (P0 + 81920 + 256 * 100 * N) + B * Dl, "note": "SURVEY 8d: 28 B/param + frames",
                              "floor_ms": (28.0 * (P0 + 81920 + 256 * 100 * N) + B * Dl) / (pk["hbm_gbs"] * 1e9) * 1e3,
                              "frac": (28.0 * (P0 + 81920 + 256 * 100 * N) + B * Dl) / (pk["hbm_gbs"] * 1e9) * 1e3 / ms}}

    # e2e: per step pinned host uint8 frames + targets -> device on a copy stream (double buffered so the copy of
    # batch i+1 overlaps the compute of batch i), loss -> host every step like src/trainer/base.py:154
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty((B, Dl), dtype=torch.uint8, device=dev), torch.empty((B, 100, N), dtype=torch.float32, device=dev))
             for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [None, None]

    def stage(i):
        fr, ap = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            if free[i % 2] is not None:
                copy_stream.wait_event(free[i % 2])            # the step that last read this slot has finished
            fr.copy_(frames_h[i % nbuf], non_blocking=True)
            ap.copy_(ap_h[i % nbuf], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        stage(0)
        last = None
        for i in range(n):
            fr, ap = slots[i % 2]
            cur.wait_event(ready[i % 2])
            loss = train_step(fr, ap, opt)
            sched.step()
            free[i % 2] = torch.cuda.Event()
            free[i % 2].record(cur)
            if i + 1 < n:
                stage(i + 1)
            last = float(loss)
        return last

    e2e_loop(2)
    torch.cuda.synchronize(); barrier(world)
    t0 = time.perf_counter()
    e2e_loop(steps)
    torch.cuda.synchronize(); barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev) / steps
    e2e = {"value": B * FRAMES_PER_TRIAL / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(B * Dl + 4 * B * 100 * N),
           "d2h_bytes_per_step": 8, "ms_per_step": e2e_s * 1e3, "statistic": f"mean of {steps} consecutive steps (max over ranks)",
           "path": "Linear.fused_train_step(pinned uint8 frames -> device copy stream) + float(loss) every step"}
    del model, opt, frames_d, ap_d, slots
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_linear_sample(B, D, N, steps=2)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle Trainer.step (torch fp32 CPU port of the reference step), B={B}, D={D}, N={N}, 2 timed steps ({dt:.2f} s each)"}
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "tf32 fwd / f32 update",
            "data": "synthetic", "config": linear_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "loss": float(loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "rrr", "linear"])
    ap.add_argument("--trials", type=int, default=400)
    ap.add_argument("--trials-test", dest="trials_test", type=int, default=80)
    ap.add_argument("--features", type=int, default=110 * 166)
    ap.add_argument("--neurons", type=int, default=144)
    ap.add_argument("--mode", default=None, choices=["exact", "dense", "classic"],
                    help="rrr operand mode (default: VS_RRR_MODE or exact; --planes implies classic)")
    ap.add_argument("--planes", type=int, default=None, help="rrr classic mode: residual planes of the 16-bit operands")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--input-dim", dest="input_dim", type=int, default=120 * 128 * 128)
    ap.add_argument("--linear-steps", dest="linear_steps", type=int, default=50)
    ap.add_argument("--cpu-budget", dest="cpu_budget", type=float, default=20.0, help="seconds of CPU work of the cpu_baseline sample")
    ap.add_argument("--ref-budget", dest="ref_budget", type=float, default=150.0, help="--impl reference: seconds for all warm-up + timed fits")
    ap.add_argument("--dropin-e2e", dest="dropin_e2e", type=int, default=2, help="samples of the float64-numpy drop-in e2e (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--strong", action="store_true", help="rrr, N > 1: ONE session with its trials sharded over the ranks (gradient all-reduce per evaluation)")
    ap.add_argument("--independent", action="store_true", help="rrr, N > 1: independent per-session fits instead of the joint shared-V model")
    ap.add_argument("--joint", action="store_true", help="(default for N > 1; kept for compatibility)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":                               # CPU arm: rank 0 alone works, no process group
        reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    line = None
    if args.workload in ("both", "rrr"):
        line = run_rrr(args, rank, world, local)
    if args.workload in ("both", "linear"):
        lin_steps = args.linear_steps if args.workload == "both" else max(args.steps, 10)
        lin = run_linear(args, rank, world, local, lin_steps)
        if args.workload == "linear":
            line = lin
        elif line is not None:
            line["linear"] = lin
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
