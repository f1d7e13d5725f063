#!/usr/bin/env python
"""bench.py -- training video frames/s of the video->spike hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload rrr|linear] [--impl reference]

Workloads (config.workload):
  rrr     BASELINE configs[1] (default): RRR rank-3, ONE session, bf16 operands, K=400 train trials of
          120 frames x (110 x 166) whisker ROI, N=144 neurons, l2=100.  A "step" is one full fit
          (src/model/rrr.py:164-202: one LBFGS.step = 20 closure evaluations + the validation pass).
          frames/s = K*120 / fit time.  N>1 GPUs: one independent session per rank (train_rrr.py:179-187
          fits sessions independently: no data-path collective), weak scaling.  --joint (configs[2]): the N
          sessions form ONE model with a shared V (rrr.py:37-49); [dV, loss] and the L-BFGS inner products
          are all-reduced over NCCL.
  linear  BASELINE configs[0]/[3]: `Linear` MLP train step (src/trainer/base.py:147-154), B=16,
          D=120*128*128, N=144: frames/s = B*120 / step time.  N>1: row-parallel first layer (each rank owns
          1/N of the pixels of W0 and of every frame; one 16 KB all-reduce per step), strong scaling.

`value`  : device-timed (CUDA events), inputs resident in HBM.
`e2e`    : same metric through the public Python API from PINNED HOST buffers, H2D/D2H inside the timed region.
`roofline`: dominant kernel, event-bracketed inside the timed region (vs_profile_*), vs MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (CPU port of the reference) timed on the host cores on a bounded sample.
--impl reference: only that CPU arm, printed in the same JSON shape.
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
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as fh:
            ent = json.load(fh).get(key)
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
def cpu_rrr_sample(K_s, Kt_s, F, N, evals=2, seed=0):
    """Host-core baseline: the oracle's autograd transcription of the reference closure, `evals` of the
    fit's 20 closure evaluations at full C and N on K_s trials.  frames/s = K_s*120 / (20 * mean eval time)."""
    from oracle import rrr_oracle as ro
    torch.set_num_threads(os.cpu_count() or 1)
    ftr, ctr, fte, cte = rrr_inputs(K_s, Kt_s, F, N, seed, pinned=False)
    data, _ = ro.preprocess_session([ftr.numpy(), fte.numpy()], [ctr.numpy().astype(np.float64), cte.numpy().astype(np.float64)],
                                    sorted_idx_42())
    td = {"s": data}
    params = ro.rrr_init(td, 3)
    ro.loss_and_grad_autograd(params, td, 100.0)             # warm-up (allocator, threads)
    t0 = time.perf_counter()
    for _ in range(evals):
        ro.loss_and_grad_autograd(params, td, 100.0)
    dt = (time.perf_counter() - t0) / evals
    fit_s = 20.0 * dt
    return K_s * FRAMES_PER_TRIAL / fit_s, fit_s, dt


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
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "rrr":
        v, fit_s, ev_s = cpu_rrr_sample(args.cpu_trials, 8, args.features, args.neurons, evals=max(1, args.steps))
        sample = (f"oracle autograd closure (CPU port of src/model/rrr.py:165-175, torch fp64), {args.cpu_trials} trials x 120 frames, "
                  f"C={args.features + 1}, N={args.neurons}: {max(1, args.steps)} timed closure evals, fit = 20 evals ({ev_s:.2f} s/eval)")
        cfg = rrr_config(args, world)
        ms = fit_s * 1e3
    else:
        v, dt = cpu_linear_sample(args.batch, args.input_dim, args.neurons, steps=max(1, min(args.steps, 3)))
        sample = f"oracle Trainer.step (CPU port of src/trainer/base.py:147-154, torch fp32), B={args.batch}, D={args.input_dim}, N={args.neurons}"
        cfg = linear_config(args, world)
        ms = dt * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.workload == "rrr" else "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- configs
def rrr_config(args, world):
    return {"workload": "rrr_joint_fit, shared V across sessions (BASELINE configs[2])" if args.joint else "rrr_single_session_fit (BASELINE configs[1])", "trials_train": args.trials, "trials_test": args.trials_test,
            "frames_per_trial": FRAMES_PER_TRIAL, "frame_shape": "1x110x166", "features": args.features, "time_bins": 100,
            "neurons": args.neurons, "rank": 3, "l2": 100, "operand_planes": args.planes,
            "operand_format": (os.environ.get("VS_RRR_OPERAND") or "bf16") + " (16-bit tensor-core operands, fp32 accumulate in TMEM)", "lbfgs": ("1 step, max_iter 20 (20 closure evals), history float32" if args.planes == 1 else "1 step, max_iter 20 (20 closure evals), history float64")
                     + (", host-driven with sharded inner products" if args.joint else ", device-driven" if os.environ.get("VS_LBFGS_DEVICE", "1") != "0" else ", host-driven"),
            "sessions": world,
            "parallelism": (f"joint model over {world} sessions, one per GPU: shared V, [dV, loss] and L-BFGS inner products all-reduced (NCCL)"
                            if args.joint else (f"independent sessions, one per GPU x{world} (no collective)" if world > 1 else "single GPU")),
            "l2_cache": "operands (1.46 GB per pass) exceed the 126 MB L2; no flush needed"}


def linear_config(args, world):
    return {"workload": "linear_mlp_train_step (BASELINE configs[0] on B200)", "batch": args.batch, "input_dim": args.input_dim,
            "neurons": args.neurons, "optimizer": "AdamW+OneCycleLR",
            "parallelism": f"first layer row-parallel over {world} GPUs (pixel shards), one 16 KB all-reduce per step" if world > 1 else "single GPU",
            "l2_cache": "weights+Adam state (6 GB) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- RRR workload
def run_rrr(args, rank, world, local):
    import vsb200 as vs
    from model.rrr import RRRGD, pack_session_from_frames, train_model, train_model_from_frames
    from optim import FusedLBFGS
    vs.require_b200()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    K, Kt, F, N = args.trials, args.trials_test, args.features, args.neurons
    sidx = sorted_idx_42()
    ftr, ctr, fte, cte = rrr_inputs(K, Kt, F, N, seed=rank, pinned=True)
    # bytes that cross PCIe per fit: the 100 selected frames of every trial (vs_h2d_select_frames) + the spike counts
    h2d = (ftr.numel() + fte.numel()) // ftr.shape[1] * len(sidx) + 4 * (ctr.numel() + cte.numel())

    # ---- resident-input measurement: operands packed once, fit repeated
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=args.planes, device=dev)
    # --joint (BASELINE configs[2]): ONE model over all ranks' sessions with a shared V (src/model/rrr.py:37-49); rank r holds
    # session r's U, b and operands, [dV, loss] and the L-BFGS inner products are all-reduced over NCCL (parallel.py)
    joint = bool(args.joint)
    eid = f"s{rank:02d}" if joint else "s"
    plan = [(f"s{r:02d}", N, F + 1, 100) for r in range(world)] if joint else None
    td = {eid: entry}
    model = RRRGD(td, 3, l2=100.0, planes=args.planes, init_plan=plan)
    model.to(dev)
    model_fmt = model.fmt
    init = {k: v.detach().clone() for k, v in model.model.items()}
    if joint:
        from parallel import train_joint_model

    def one_fit():
        with torch.no_grad():
            for k, v in init.items():
                model.model[k].copy_(v)
        if joint:
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
    vs.lib.vs_profile_enable(0)
    evals = (model.n_closure_evals - evals0) / args.steps
    value = world * K * FRAMES_PER_TRIAL / (ms * 1e-3)

    # roofline of the dominant kernel (the forward tcgen05 GEMM, Z = X U): algorithmic FLOPs (2*K*T*C*N per
    # contraction, i.e. the dense formulation) summed over the launches of the timed region / their summed duration.
    # With VS_RRR_DENSE=0 the factorised backward GEMM runs under the same tag and is counted the same way.
    C = F + 1
    Np16 = (N + 15) // 16 * 16
    plane_passes = 1 if args.planes == 1 else (3 if args.planes == 2 else 6)
    n_contr = 1 if n_bwd > 0 else 2                             # contractions per closure evaluation under tag 0
    algo_flops = args.steps * (evals * n_contr * (2.0 * K * 100 * C * N) + 2.0 * Kt * 100 * C * N)
    exec_flops = args.steps * (evals * n_contr * (2.0 * K * 100 * F * 3 * Np16) + 2.0 * Kt * 100 * F * 3 * Np16) * plane_passes
    pk, pk_kind = peaks()
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    ach = algo_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    key = f"rrr_K{K}_F{F}_N{N}_planes{args.planes}"
    roof = {"bound": "tensor", "kernel": "vs::tc::gemm_tn_pair_kernel (forward Z = X U; tcgen05 cta_group::2 kind::f16, UMMA 256xN over a CTA pair)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
            "traffic": ncu_traffic(key), "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01_ncu_traffic.json)",
            "peak_source": f"{pk_kind} bf16_tflops_sustained",
            "launches": n_gemm, "avg_launch_ms": gemm_ms / max(n_gemm, 1), "share_of_step": gemm_ms / (ms * args.steps),
            "executed_tflops": exec_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
            "note": "achieved counts ALGORITHMIC flops of the dense formulation; the factorised forward executes r=3x that (executed_tflops)"}
    if n_bwd > 0:
        # second kernel of the closure: the per-time-bin dense backward streams Xb once (HBM-bound at a third of the flops)
        Kp = (K + 15) // 16 * 16
        bwd_bytes = 2.0 * F * 100 * Kp + 2.0 * Np16 * 100 * Kp + 4.0 * F * 3 * Np16      # Xb + R operand + G out
        bw = bwd_bytes * n_bwd / (bwd_ms * 1e-3) / 1e9
        roof["backward"] = {"bound": "hbm", "kernel": "vs::tc::rrr_bwd_dense_pair_kernel (dU: D_t = X_t^T R_t per time bin in TMEM, rank-one updates in registers)",
                            "achieved": bw, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bw / pk["hbm_gbs"],
                            "traffic": ncu_traffic(key, "bwd_bytes_per_launch"), "launches": n_bwd, "avg_launch_ms": bwd_ms / n_bwd,
                            "share_of_step": bwd_ms / (ms * args.steps),
                            "algorithmic_tflops": 2.0 * K * 100 * C * N * n_bwd / (bwd_ms * 1e-3) / 1e12}

    # ---- parity of the timed configuration (outside every timed region): the same fit with 3 operand planes and a float64
    # L-BFGS history -- the mode the tests pin against the float64 reference to ~1e-6 -- on the same session
    parity = None
    if rank == 0 and not args.no_parity and args.planes == 1 and not joint:
        entry3 = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=3, device=dev)
        m3 = RRRGD({"s": entry3}, 3, l2=100.0, planes=3)
        m3.to(dev)
        # (1) one closure evaluation at identical (initial) parameters: loss and gradient, timed mode vs 3 planes
        gp = torch.Generator().manual_seed(1)
        with torch.no_grad():                      # perturbed start (at the exact init the gradient of b is ~0 by construction)
            for k, v in init.items():
                x = v + (0.02 * torch.randn(v.shape, generator=gp, dtype=torch.float64)).to(dev)
                model.model[k].copy_(x)
                m3.model[k].copy_(x)
        l1, l3 = float(model.loss_and_grad(td, 0)), float(m3.loss_and_grad({"s": entry3}, 0))
        gerr = {k.split("_")[-1]: float((model.model[k].grad - m3.model[k].grad).abs().max() / m3.model[k].grad.abs().max()) for k in init}
        with torch.no_grad():
            for k, v in init.items():
                m3.model[k].copy_(v)
        # (2) the whole fit (20 closure evaluations of an un-line-searched L-BFGS amplify any perturbation)
        _, r3 = train_model(m3, {"s": entry3}, m3.make_optimizer(), "tmp", save=False)
        ref_sse = float(r3["mse_val_mean"])
        parity = {"per_eval_loss_rel_diff": abs(l1 - l3) / abs(l3), "per_eval_grad_max_abs_diff_over_max_abs": gerr,
                  "fit_val_sse": float(mse), "fit_val_sse_3plane_f64hist": ref_sse, "fit_rel_diff": abs(float(mse) - ref_sse) / ref_sse,
                  "tolerance": 1e-3,
                  "note": "reference = 3 bf16 residual planes + float64 L-BFGS history (pinned to the float64 reference by the tests; "
                          "1.0e-4 from a float64 dense torch fit at this size, profiles/r01_parity_probe_noise_data.txt)"}
        del entry3, m3, r3
        torch.cuda.empty_cache()

    # ---- end to end: pinned host uint8 frames -> R0 on device -> init -> fit -> validation loss on the host
    def e2e_fit():
        if joint:
            ent = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=args.planes, device=dev)
            m = RRRGD({eid: ent}, 3, l2=100.0, planes=args.planes, init_plan=plan, device=dev)
            _, res = train_joint_model(m, {eid: ent})
        else:
            m, res, _ = train_model_from_frames(ftr, ctr, fte, cte, sidx, l2=100.0, n_comp=3, planes=args.planes)
        return float(res["mse_val_mean"])                    # device -> host read of the result

    del model, td, entry
    torch.cuda.empty_cache()
    import gc
    e2e_fit(); e2e_fit()
    gc.collect(); gc.disable()                      # no collector pauses inside the timed region (re-enabled below)
    torch.cuda.synchronize(); barrier(world)
    n_e2e = max(5, args.steps)
    each = []
    for i in range(3 * n_e2e):
        t1 = time.perf_counter()
        val = e2e_fit()
        each.append((time.perf_counter() - t1) * 1e3)
        gc.collect()                                # between fits, outside the per-fit timing: frees the previous fit's operands
        # a shared host can stall single fits by 100+ ms: keep sampling (up to 3x) until the fastest and the median agree to 25 %
        # (ranks decide together: the joint model's fits contain collectives)
        if i + 1 >= n_e2e and max_over_ranks(1.0 if float(np.median(each)) > 1.25 * min(each) else 0.0, world, dev) == 0.0:
            break
    n_e2e = len(each)
    torch.cuda.synchronize(); barrier(world)
    mean_s = max_over_ranks(float(np.mean(each)) * 1e-3, world, dev)
    # this path crosses the host 20+ times per fit (init stream threads, one sync per L-BFGS iteration, PCIe): on a shared
    # box single iterations are hit by 100+ ms of host jitter, so the headline uses the MEDIAN iteration (max over ranks);
    # the mean and every sample are reported next to it
    e2e_s = max_over_ranks(float(np.median(each)) * 1e-3, world, dev)
    gc.enable()
    e2e = {"value": world * K * FRAMES_PER_TRIAL / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
           "ms_per_step": e2e_s * 1e3, "statistic": f"median of {n_e2e} fits", "mean_ms_per_step": mean_s * 1e3,
           "ms_each_rank0": [round(x, 2) for x in each],
           "path": ("pack_session_from_frames(pinned uint8 frames) -> RRRGD(init_plan) -> parallel.train_joint_model -> float(mse_val_mean)" if joint
                    else "model.rrr.train_model_from_frames(pinned uint8 frames) -> float(mse_val_mean)")}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, fit_s, ev_s = cpu_rrr_sample(args.cpu_trials, 8, F, N, evals=2)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle autograd closure (torch fp64 CPU), {args.cpu_trials} trials x 120 frames at full C={C}, N={N}: "
                         f"2 timed closure evals ({ev_s:.2f} s each), fit = 20 evals"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if model_fmt == 1 else "bf16",
            "data": "synthetic", "config": rrr_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "closure_evals_per_step": evals, "val_sse": float(mse), "val_sse_e2e": val, "parity": parity}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- Linear workload
def run_linear(args, rank, world, local):
    import vsb200 as vs
    from tests.helpers import make_linear_model
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
    for i in range(args.warmup):
        step(i, frames_d, ap_d)
    torch.cuda.synchronize(); barrier(world)
    vs.lib.vs_launch_count_reset(); vs.lib.vs_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        loss = step(i, frames_d, ap_d)
    e1.record(); torch.cuda.synchronize(); barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / args.steps
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
            "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic(f"linear_B{B}_D{D}_N{N}") if world == 1 else None, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01_ncu_traffic.json)", "peak_source": f"{pk_kind} hbm_gbs", "launches": n_k,
            "avg_launch_ms": k_ms / max(n_k, 1), "share_of_step": k_ms / (ms * args.steps)}

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
    e2e_loop(args.steps)
    torch.cuda.synchronize(); barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev) / args.steps
    e2e = {"value": B * FRAMES_PER_TRIAL / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(B * Dl + 4 * B * 100 * N),
           "d2h_bytes_per_step": 8, "ms_per_step": e2e_s * 1e3,
           "path": "Linear.fused_train_step(pinned uint8 frames -> device copy stream) + float(loss) every step"}
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_linear_sample(B, D, N, steps=2)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle Trainer.step (torch fp32 CPU port of the reference step), B={B}, D={D}, N={N}, 2 timed steps ({dt:.2f} s each)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "tf32 fwd / f32 update",
            "data": "synthetic", "config": linear_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "loss": float(loss)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rrr", choices=["rrr", "linear"])
    ap.add_argument("--trials", type=int, default=400)
    ap.add_argument("--trials-test", dest="trials_test", type=int, default=80)
    ap.add_argument("--features", type=int, default=110 * 166)
    ap.add_argument("--neurons", type=int, default=144)
    ap.add_argument("--planes", type=int, default=1)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--input-dim", dest="input_dim", type=int, default=120 * 128 * 128)
    ap.add_argument("--cpu-trials", dest="cpu_trials", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--joint", action="store_true", help="rrr: one joint model with a shared V over all ranks' sessions (configs[2])")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.workload == "rrr" else 50
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":                               # CPU arm: rank 0 alone works, no process group
        reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    if args.workload == "rrr":
        run_rrr(args, rank, world, local)
    else:
        run_linear(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
