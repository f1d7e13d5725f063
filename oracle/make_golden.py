"""Generate tests/golden/*.npz from the UNMODIFIED reference modules under /root/reference/src.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The fixtures are small and committed; tests/test_oracle_golden.py replays them against the oracle
and the GPU tests replay them against the CUDA path.  Missing third-party modules that the
reference imports at module scope but the hot path never calls (matplotlib, cebra, torcheval) are
stubbed, as described in SURVEY.md section 8c.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub_modules():
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = lambda *a, **k: (_Any(), _Any())
    plt.colorbar = lambda *a, **k: _Any()
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    cebra = types.ModuleType("cebra")
    cebra.CEBRA = _Any
    sys.modules["cebra"] = cebra

    class R2Score:
        def reset(self):
            self.p, self.t = [], []

        def to(self, device):
            return self

        def update(self, pred, true):
            self.p.append(pred.double().flatten()); self.t.append(true.double().flatten())

        def compute(self):
            p, t = torch.cat(self.p), torch.cat(self.t)
            return 1.0 - torch.sum((t - p) ** 2) / torch.sum((t - t.mean()) ** 2)

    te, tm = types.ModuleType("torcheval"), types.ModuleType("torcheval.metrics")
    tm.R2Score = R2Score
    te.metrics = tm
    sys.modules["torcheval"], sys.modules["torcheval.metrics"] = te, tm


def golden_metrics():
    from utils.metric_utils import bits_per_spike, neg_log_likelihood
    from utils.utils import _one_hot, _std, metrics_list, set_seed
    rng = np.random.RandomState(0)
    spikes = rng.poisson(0.4, size=(5, 100, 3)).astype(float)
    rates = np.clip(0.4 + 0.1 * rng.standard_normal((5, 100, 3)), 1e-3, None)
    out = dict(spikes=spikes, rates=rates,
               nll=neg_log_likelihood(rates.copy(), spikes),
               bps=bits_per_spike(rates.copy(), spikes),
               bps_n0=bits_per_spike(rates[:, :, [0]].copy(), spikes[:, :, [0]]))
    # metrics_list with its trial-count quirk: K=4 trials, N=6 neurons, gt passed as (N,T,K)
    rng = np.random.RandomState(1)
    gt = torch.from_numpy(rng.poisson(0.5, size=(4, 100, 6)).astype(np.float32))
    pred = torch.from_numpy(np.clip(0.5 + 0.2 * rng.standard_normal((4, 100, 6)), 1e-3, None).astype(np.float32))
    res = metrics_list(gt=gt.transpose(-1, 0), pred=pred.transpose(-1, 0), metrics=["bps", "rsquared"], device="cpu")
    out.update(ml_gt=gt.numpy(), ml_pred=pred.numpy(), ml_bps=res["bps"], ml_rsquared=res["rsquared"])
    # z-score + one-hot helpers
    arr = rng.standard_normal((7, 5, 3))
    arr[:, 2, 1] = 4.0   # constant column -> std clipped to 1e-8
    z, mean, std = _std(arr)
    out.update(std_in=arr, std_z=z, std_mean=mean, std_std=std)
    ch = rng.choice([-1.0, 1.0], size=(6, 1))
    out.update(oh_in=ch, oh_out=_one_hot(ch, 4))
    # frame selection (train_rrr.py:41,48-49)
    set_seed(42)
    idx = np.sort(np.random.choice(119, 100, replace=False))
    out.update(sorted_idx=idx.astype(np.int64),
               sorted_idx_sha256=np.frombuffer(hashlib.sha256(idx.astype(np.int64).tobytes()).digest(), dtype=np.uint8))
    np.savez_compressed(os.path.join(OUT, "metrics_kat.npz"), **out)
    print("metrics_kat", out["nll"], out["bps"], out["ml_bps"], out["ml_rsquared"])


def golden_linear():
    """Reference Linear + AdamW + OneCycleLR + PoissonNLLLoss for 4 steps on a tiny frame size."""
    from model.linear import Linear
    from utils.config_utils import config_from_kwargs, update_config
    from utils.utils import set_seed
    H = W = 8
    N, B, T_f = 6, 4, 120
    D = T_f * H * W
    config = config_from_kwargs({"model": "include:/root/reference/config/model/linear_video.yaml"})
    config = update_config("/root/reference/config/train/linear_video.yaml", config)
    set_seed(config.seed)
    config["model"]["encoder"]["input_dim"] = D
    config["model"]["decoder"]["output_dim"] = 100 * N
    model = Linear(config.model)
    init_sd = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(model.parameters(), lr=config.optimizer.lr, weight_decay=config.optimizer.wd, eps=config.optimizer.eps)
    sched = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, total_steps=40, max_lr=config.optimizer.lr,
                                                pct_start=config.optimizer.warmup_pct, div_factor=config.optimizer.div_factor)
    crit = torch.nn.PoissonNLLLoss(reduction="none", log_input=True)
    g = torch.Generator().manual_seed(0)
    frames = (torch.randint(1, 9, (6, B, T_f, 1, H, W), generator=g, dtype=torch.uint8)
              * (torch.rand((6, B, T_f, 1, H, W), generator=g) < 0.05).to(torch.uint8))
    ap = torch.poisson(torch.full((6, B, 100, N), 0.3), generator=g)
    losses, lrs, b1s = [], [], []
    first_logits = None
    for s in range(6):
        lrs.append(opt.param_groups[0]["lr"]); b1s.append(opt.param_groups[0]["betas"][0])
        x = frames[s].float()                                  # loader/base.py:39
        out = model(torch.cat([x.flatten(1)], dim=-1))          # trainer/base.py:64-70
        if first_logits is None:
            first_logits = out.detach().clone().numpy()
        loss = crit(out, ap[s]).mean()                          # trainer/base.py:141-143
        loss.backward()
        opt.step(); sched.step(); opt.zero_grad()
        losses.append(loss.item())
    model.eval()
    with torch.no_grad():
        rates = torch.exp(model(frames[0].float().flatten(1))).numpy()
    final_sd = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "linear_small.npz"), frames=frames.numpy(), ap=ap.numpy(), losses=np.array(losses),
                        lrs=np.array(lrs), beta1s=np.array(b1s), first_logits=first_logits, final_rates=rates,
                        H=H, W=W, N=N, B=B, total_steps=40,
                        **{"init/" + k: v for k, v in init_sd.items() if "layers.0.weight" not in k or "decoder" in k},
                        **{"final/" + k: v for k, v in final_sd.items() if k != "encoder.layers.0.weight"})
    # the 256 x D first-layer init is reproducible from the seed; keep only a checksum + a slice
    w0 = init_sd["encoder.layers.0.weight"]
    w0f = final_sd["encoder.layers.0.weight"]
    np.savez_compressed(os.path.join(OUT, "linear_small_w0.npz"), w0_slice=w0[:4, :64], w0_sum=np.float64(w0.astype(np.float64).sum()),
                        w0_abs=np.float64(np.abs(w0.astype(np.float64)).sum()),
                        w0_final_rows=w0f[::37, ::13].copy(), w0_delta_sum=np.float64((w0f.astype(np.float64) - w0).sum()),
                        w0_delta_abs=np.float64(np.abs(w0f.astype(np.float64) - w0).sum()))
    print("linear_small losses", losses, "lrs", lrs[:3], "b1", b1s[:3])
    # OneCycleLR table for the reference's own schedule length (25 batches x 200 epochs)
    p = torch.nn.Parameter(torch.zeros(1))
    o = torch.optim.AdamW([p], lr=5e-5, weight_decay=0.01, eps=1e-8)
    sc = torch.optim.lr_scheduler.OneCycleLR(o, total_steps=5000, max_lr=5e-5, pct_start=0.15, div_factor=10)
    tab = []
    for s in range(5000):
        if s in (0, 1, 2, 375, 748, 749, 750, 751, 2500, 4998, 4999):
            tab.append((s, o.param_groups[0]["lr"], o.param_groups[0]["betas"][0]))
        o.step()
        if s < 4999:
            sc.step()
    np.savez_compressed(os.path.join(OUT, "onecycle.npz"), table=np.array(tab, dtype=np.float64))


def _rrr_problem(seed, K, Kt, F, N, signal=True):
    rng = np.random.default_rng(seed)
    Xtr = rng.integers(0, 256, size=(K, 120, F)).astype(np.uint8)
    Xte = rng.integers(0, 256, size=(Kt, 120, F)).astype(np.uint8)
    Wt = rng.standard_normal((F, 3)) / np.sqrt(F)
    Vt = rng.standard_normal((3, 120))
    A = rng.standard_normal((3, N))

    def rates(X):
        z = ((X.astype(float) - 127.5) / 74.0) @ Wt
        lat = np.einsum("ktj,jt->ktj", z, Vt)[:, :100]
        return np.exp(0.6 * np.einsum("ktj,jn->ktn", lat, A) - 1.0)

    ytr = rng.poisson(rates(Xtr)).astype(np.float64)
    yte = rng.poisson(rates(Xte)).astype(np.float64)
    return Xtr, Xte, ytr, yte


def golden_rrr():
    """Reference RRRGD / train_model_main on a small planted problem, driven through the same
    preprocessing steps as src/train_rrr.py:108-171 (transcribed call by call)."""
    from scipy.ndimage import gaussian_filter1d
    from model.rrr import train_model_main
    from utils.utils import _std, set_seed
    set_seed(42)
    sorted_idx = np.sort(np.random.choice(119, 100, replace=False))
    out = {}
    for name, (seed, K, Kt, F, N) in {"a": (0, 40, 12, 96, 10), "b": (1, 24, 8, 50, 7)}.items():
        Xtr, Xte, ytr, yte = _rrr_problem(seed, K, Kt, F, N)
        td = {"e1": {"X": [Xtr.astype(np.float64), Xte.astype(np.float64)], "y": [ytr.copy(), yte.copy()], "setup": {}}}
        gt = td["e1"]["y"][1]
        for i in range(2):
            td["e1"]["y"][i] = gaussian_filter1d(td["e1"]["y"][i], 2, axis=1)
        _, mean_X, std_X = _std(td["e1"]["X"][0])
        _, mean_y, std_y = _std(td["e1"]["y"][0])
        for i in range(2):
            Kk, Tt = td["e1"]["X"][i].shape[:2]
            td["e1"]["X"][i] = (td["e1"]["X"][i] - mean_X) / std_X
            td["e1"]["X"][i] = np.concatenate([td["e1"]["X"][i], np.ones((Kk, Tt, 1))], axis=2)
            td["e1"]["X"][i] = td["e1"]["X"][i][:, sorted_idx]
            td["e1"]["y"][i] = (td["e1"]["y"][i] - mean_y) / std_y
        td["e1"]["setup"].update(mean_X_Tv=mean_X, std_X_Tv=std_X, mean_y_TN=mean_y, std_y_TN=std_y)
        model, mse_val = train_model_main(td, l2=100, n_comp=3, model_fname="/tmp/_golden_rrr", save=False)
        with torch.no_grad():
            _, y_fr, pred_fr = model.predict_y_fr(td, "e1", 1)
            # loss at the FINAL parameters (one more closure evaluation, no step)
            mses = model.compute_MSE_RRRGD(td, 0)
            reg = model.regression_loss()
            final_loss = float(mses["e1"].sum() + reg["e1"])
        out.update({f"{name}/Xtr": Xtr, f"{name}/Xte": Xte, f"{name}/ytr": ytr, f"{name}/yte": yte,
                    f"{name}/X0_proc": td["e1"]["X"][0][:3], f"{name}/y0_proc": td["e1"]["y"][0][:3],
                    f"{name}/U": model.model["e1_U"].detach().numpy(), f"{name}/V": model.model["V"].detach().numpy(),
                    f"{name}/b": model.model["e1_b"].detach().numpy(),
                    f"{name}/mses_val": mse_val["mses_val"]["e1"].detach().numpy(),
                    f"{name}/mse_val_mean": float(mse_val["mse_val_mean"]), f"{name}/final_train_loss": final_loss,
                    f"{name}/pred_fr": pred_fr.numpy(), f"{name}/gt": gt})
        print("rrr", name, float(mse_val["mse_val_mean"]), final_loss)
    # init KAT: N=2, ncoef-1=3, r=3, T=100
    np.random.seed(0)
    U = np.random.normal(size=(2, 3, 3)) / np.sqrt(100 * 3)
    out["init_U_first4"] = U.ravel()[:4]
    out["sorted_idx"] = sorted_idx
    np.savez_compressed(os.path.join(OUT, "rrr_small.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    _stub_modules()
    sys.path.insert(0, REF)
    torch.set_num_threads(8)
    golden_metrics()
    golden_linear()
    golden_rrr()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
