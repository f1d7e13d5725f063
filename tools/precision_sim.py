"""CPU float64 emulation of operand-precision choices for the RRR closure (design aid, not product code).

Runs the reference fit (src/model/rrr.py:164-202: one un-line-searched LBFGS.step, 20 closure evaluations) on a
synthetic session of bench.py's distribution with selected operands ROUNDED to a given number of significant bits
(accumulation stays float64) and prints the deviation of the final validation SSE from the all-float64 fit.
It answers "which operand needs how many bits for the whole fit to land within 1e-3" before a kernel is written.

    python tools/precision_sim.py [F=4000] [K=400] [N=144]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from scipy.ndimage import gaussian_filter1d
import bench

torch.set_num_threads(os.cpu_count() or 1)
DEV = torch.device("cuda" if torch.cuda.is_available() else "cpu")      # float64 on the GPU: a full-size fit takes ~1 s
F = int(os.environ.get("F", 4000)); K = int(os.environ.get("K", 400)); N = int(os.environ.get("N", 144)); Kt = 80
T, r, l2 = 100, 3, 100.0
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, 0, pinned=False, signal=os.environ.get("SIGNAL", "0") == "1")
sidx = torch.as_tensor(bench.sorted_idx_42())
GEN = torch.Generator(device=DEV); GEN.manual_seed(1234)


def rnd(x, bits):
    """round to `bits` significant bits (bits=None: exact)"""
    if bits is None:
        return x
    m, e = torch.frexp(x)
    return torch.ldexp(torch.round(m * (1 << bits)) / (1 << bits), e)


def split2(x, bits):
    """hi + lo residual planes, each `bits` significant bits"""
    hi = rnd(x, bits)
    return hi + rnd(x - hi, bits)


# ---- R0 (train_rrr.py:108-171): per-(frame, feature) z-score with train statistics, 100 selected frames, time-major
Xi = ftr[:, sidx].to(DEV).double().permute(1, 0, 2).contiguous()            # (T, K, F) integers
Xie = fte[:, sidx].to(DEV).double().permute(1, 0, 2).contiguous()
mu = Xi.mean(1); sd = Xi.std(1, unbiased=False).clamp_min(1e-8); isd = 1.0 / sd          # (T, F)
m_int = torch.round(mu); delta = mu - m_int
Xc, Xce = Xi - m_int[:, None], Xie - m_int[:, None]                   # exact integer operands |x| <= 255
Xn, Xne = (Xi - mu[:, None]) * isd[:, None], (Xie - mu[:, None]) * isd[:, None]
ytr = gaussian_filter1d(ctr.numpy().astype(np.float64), 2, axis=1); yte = gaussian_filter1d(cte.numpy().astype(np.float64), 2, axis=1)
my, sy = ytr.mean(0), np.clip(ytr.std(0), 1e-8, None)
y = torch.from_numpy((ytr - my) / sy).to(DEV).permute(1, 0, 2).contiguous()   # (T, K, N)
ye = torch.from_numpy((yte - my) / sy).to(DEV).permute(1, 0, 2).contiguous()
np.random.seed(0)
U0 = torch.from_numpy(np.random.normal(size=(N, F, r)) / np.sqrt(T * r)).to(DEV); V0 = torch.from_numpy(np.random.normal(size=(r, T)) / np.sqrt(T * r)).to(DEV)
b0 = y.mean(1).T.contiguous()                                         # (N, T)
del ftr, fte


def closure(U, V, b, cfg):
    """loss and gradients under the operand-rounding model cfg"""
    fwd = cfg.get("fwd", "exact")
    beta = torch.einsum("ncj,jt->tcn", U, V)                          # (T, F, N)
    if fwd == "exact":
        yd = torch.bmm(Xn, beta)
    elif fwd == "dense_exactA":                                        # integer A exact, B = beta/sigma as hi(+lo) planes
        bq = beta * isd[:, :, None]
        bq = split2(bq, cfg["b_bits"]) if cfg.get("b_planes", 2) == 2 else rnd(bq, cfg["b_bits"])
        yd = torch.bmm(Xc, bq) - torch.einsum("tc,tcn->tn", delta, bq)[:, None, :]
    elif fwd == "fact":                                                # factorised: z-scored A and U rounded, one plane each
        Xq = cfg["_Xq"]; Uq = rnd(U, cfg["b_bits"])
        Z = torch.matmul(Xq.reshape(T * K, F), Uq.permute(1, 0, 2).reshape(F, N * r)).reshape(T, K, N, r)
        yd = torch.einsum("tknj,jt->tkn", Z, V)
    yd = yd * (1.0 - cfg.get("shrink", 0.0))
    if cfg.get("fwd_noise"):
        yd = yd * (1.0 + cfg["fwd_noise"] * torch.randn(yd.shape, generator=GEN, device=DEV, dtype=torch.float64))
    R = yd + b.T[:, None, :] - y
    loss = (R ** 2).sum() + l2 * ((beta ** 2).sum() + (b ** 2).sum())
    # backward
    bwd = cfg.get("bwd", "exact")
    if bwd == "exact":
        D = torch.bmm(Xn.transpose(1, 2), R)                          # (T, F, N) = dbeta_t / 2 without the penalty
    elif bwd == "dense_exactA":
        Rq = split2(R, cfg["r_bits"]) if cfg.get("r_planes", 2) == 2 else rnd(R, cfg["r_bits"])
        D = (torch.bmm(Xc.transpose(1, 2), Rq) - delta[:, :, None] * Rq.sum(1)[:, None, :]) * isd[:, :, None]
    elif bwd == "dense_q":                                             # z-scored A rounded, R rounded, one plane each (today's default)
        D = torch.bmm(cfg["_Xq"].transpose(1, 2), rnd(R, cfg["r_bits"]))
    D = D * (1.0 - cfg.get("shrink_b", 0.0))
    if cfg.get("grad_noise"):
        D = D * (1.0 + cfg["grad_noise"] * torch.randn(D.shape, generator=GEN, device=DEV, dtype=torch.float64))
    dbeta = 2.0 * D + 2.0 * l2 * beta
    dU = torch.einsum("tcn,jt->ncj", dbeta, V)
    db = 2.0 * R.sum(1).T + 2.0 * l2 * b
    dvs = cfg.get("dv", "exact")
    if dvs == "exact":
        dV = torch.einsum("ncj,tcn->jt", U, dbeta)
        if cfg.get("dv_noise"):
            dV = dV * (1.0 + cfg["dv_noise"] * torch.randn(dV.shape, generator=GEN, device=DEV, dtype=torch.float64))
    elif dvs == "D_Uq":                                                # backward epilogue with U kept at u_bits
        dV = torch.einsum("ncj,tcn->jt", rnd(U, cfg["u_bits"]), 2.0 * D) + 2.0 * l2 * torch.einsum("ncj,tcn->jt", U, beta)
    elif dvs == "Z":                                                   # from a factorised product with rounded operands
        Xq = cfg["_Xq_dv"]; Uq = rnd(U, cfg["dv_b_bits"])
        Z = torch.matmul(Xq.reshape(T * K, F), Uq.permute(1, 0, 2).reshape(F, N * r)).reshape(T, K, N, r)
        dV = 2.0 * torch.einsum("tkn,tknj->jt", R, Z) + 2.0 * l2 * torch.einsum("ncj,tcn->jt", U, beta)
    return loss, dU, dV, db


def lbfgs_fit(cfg):
    """torch.optim.LBFGS.step semantics (oracle/rrr_oracle.py lbfgs_step) on the flat vector [U, b, V]"""
    from oracle.rrr_oracle import lbfgs_step
    if cfg.get("a_bits"):
        cfg["_Xq"] = rnd(Xn, cfg["a_bits"])
    if cfg.get("dv") == "Z":
        cfg["_Xq_dv"] = rnd(Xn, cfg["dv_a_bits"])
    nU, nb = N * F * r, N * T
    trace = []

    def cl(x):
        xt = torch.from_numpy(x).to(DEV)
        U, b, V = xt[:nU].reshape(N, F, r), xt[nU:nU + nb].reshape(N, T), xt[nU + nb:].reshape(r, T)
        loss, dU, dV, db = closure(U, V, b, cfg)
        trace.append(float(loss))
        return float(loss), torch.cat([dU.reshape(-1), db.reshape(-1), dV.reshape(-1)]).cpu().numpy()

    x0 = torch.cat([U0.reshape(-1), b0.reshape(-1), V0.reshape(-1)]).cpu().numpy()
    x, _ = lbfgs_step(cl, x0)
    xt = torch.from_numpy(x).to(DEV)
    U, b, V = xt[:nU].reshape(N, F, r), xt[nU:nU + nb].reshape(N, T), xt[nU + nb:].reshape(r, T)
    beta = torch.einsum("ncj,jt->tcn", U, V)
    val = float(((torch.bmm(Xne, beta) + b.T[:, None, :] - ye) ** 2).sum())
    return val, trace


CONFIGS = {
    "exact": {},
    # today's default: z-scored bf16 A, bf16 U, bf16 R, dV from the same Z
    "bf16_1plane": {"fwd": "fact", "a_bits": 8, "b_bits": 8, "bwd": "dense_q", "r_bits": 8, "dv": "Z", "dv_a_bits": 8, "dv_b_bits": 8},
    "f16_1plane": {"fwd": "fact", "a_bits": 11, "b_bits": 11, "bwd": "dense_q", "r_bits": 11, "dv": "Z", "dv_a_bits": 11, "dv_b_bits": 11},
    # proposed: integer A exact, small operand hi/lo
    "exactA_bf16x2": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8},
    "exactA_f16x2": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11},
    "exactA_bf16x2_shrink6e-5": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "shrink": 6e-5},
    "exactA_bf16x2_shrink1e-5": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "shrink": 1e-5},
    "exactA_bf16x2_dvU11": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "dv": "D_Uq", "u_bits": 11},
    "exactA_bf16x2_dvU8": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "dv": "D_Uq", "u_bits": 8},
    "exactA_bf16x2_dvZ8": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "dv": "Z", "dv_a_bits": 8, "dv_b_bits": 8},
    "exactA_bf16x2_dvZ11": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "dv": "Z", "dv_a_bits": 11, "dv_b_bits": 11},
    # which single operand matters: exact everywhere except one
    "only_fwd_bf16": {"fwd": "fact", "a_bits": 8, "b_bits": 8},
    "only_fwd_b1plane8": {"fwd": "dense_exactA", "b_bits": 8, "b_planes": 1},
    "only_bwd_r8": {"bwd": "dense_exactA", "r_bits": 8, "r_planes": 1},
    "only_bwd_q8": {"bwd": "dense_q", "a_bits": 8, "r_bits": 8},
    "only_dvZ8": {"dv": "Z", "dv_a_bits": 8, "dv_b_bits": 8},
    "only_shrink6e-5": {"shrink": 6e-5},
    "only_shrink1e-5": {"shrink": 1e-5},
    "only_shrink_b6e-5": {"shrink_b": 6e-5},
    # tolerance curves: iid relative noise on one quantity, everything else exact
    "fwd_noise1e-3": {"fwd_noise": 1e-3}, "fwd_noise1e-4": {"fwd_noise": 1e-4}, "fwd_noise1e-5": {"fwd_noise": 1e-5}, "fwd_noise1e-6": {"fwd_noise": 1e-6},
    "grad_noise1e-3": {"grad_noise": 1e-3}, "grad_noise1e-4": {"grad_noise": 1e-4}, "grad_noise1e-5": {"grad_noise": 1e-5}, "grad_noise1e-6": {"grad_noise": 1e-6},
    "dv_noise1e-2": {"dv_noise": 1e-2}, "dv_noise1e-3": {"dv_noise": 1e-3}, "dv_noise1e-4": {"dv_noise": 1e-4}, "dv_noise1e-5": {"dv_noise": 1e-5},
    "exactA_bf16x2_dvnoise1e-3": {"fwd": "dense_exactA", "b_bits": 8, "bwd": "dense_exactA", "r_bits": 8, "dv_noise": 1e-3},
    "exactA_f16x2_shrink6e-5": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "shrink": 6e-5},
    "exactA_f16x2_shrink1e-5": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "shrink": 1e-5},
    "exactA_f16x2_dvU11": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "dv": "D_Uq", "u_bits": 11},
    "exactA_f16x2_dvZ11": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "dv": "Z", "dv_a_bits": 11, "dv_b_bits": 11},
    "exactA_f16x2_dvZ8": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "dv": "Z", "dv_a_bits": 8, "dv_b_bits": 8},
    # exact-A forward with ONE B plane (cheapest dense forward), hi/lo residual in the backward
    "exactA_fwd1x11_bwd2x11": {"fwd": "dense_exactA", "b_bits": 11, "b_planes": 1, "bwd": "dense_exactA", "r_bits": 11},
    "exactA_fwd2x11_bwd1x11": {"fwd": "dense_exactA", "b_bits": 11, "bwd": "dense_exactA", "r_bits": 11, "r_planes": 1},
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    t0 = time.time()
    ref_val, ref_tr = lbfgs_fit(dict(CONFIGS["exact"]))
    print(f"F={F} K={K} N={N}: exact fit val SSE {ref_val:.6f} evals {len(ref_tr)} ({time.time() - t0:.0f} s)", flush=True)
    for nm in names:
        if nm == "exact":
            continue
        t0 = time.time()
        val, tr = lbfgs_fit(dict(CONFIGS[nm]))
        dev = max(abs(a - b) / abs(b) for a, b in zip(tr, ref_tr))
        print(f"{nm:28s} val SSE rel diff {abs(val - ref_val) / ref_val:.3e} | eval-0 loss rel diff {abs(tr[0] - ref_tr[0]) / ref_tr[0]:.3e} "
              f"max over evals {dev:.3e} ({time.time() - t0:.0f} s)", flush=True)
