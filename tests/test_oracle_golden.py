"""CPU: pin the oracle (oracle/*.py) against fixtures produced by the reference's own modules
(oracle/make_golden.py).  If these fail the oracle is wrong and no GPU parity claim stands."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import linear_oracle as lo
from oracle import metrics_oracle as mo
from oracle import rrr_oracle as ro


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ----------------------------------------------------------------------------- metrics / helpers
def test_bits_per_spike_kat(golden_dir):
    g = _load(golden_dir, "metrics_kat.npz")
    assert mo.neg_log_likelihood(g["rates"], g["spikes"]) == pytest.approx(float(g["nll"]), rel=1e-14)
    assert float(g["nll"]) == pytest.approx(1264.203734444004, rel=1e-14)
    assert mo.bits_per_spike(g["rates"], g["spikes"]) == pytest.approx(float(g["bps"]), rel=1e-12)
    assert mo.bits_per_spike(g["rates"][:, :, [0]], g["spikes"][:, :, [0]]) == pytest.approx(float(g["bps_n0"]), rel=1e-12)
    # the null model scores exactly 0
    null = np.tile(g["spikes"].mean(axis=(0, 1), keepdims=True), (5, 100, 1))
    assert mo.bits_per_spike(null, g["spikes"]) == 0.0


def test_metrics_list_quirk(golden_dir):
    g = _load(golden_dir, "metrics_kat.npz")
    res = mo.metrics_list(g["ml_gt"].astype(np.float64), g["ml_pred"].astype(np.float64))
    assert res["bps"] == pytest.approx(float(g["ml_bps"]), rel=1e-6)
    assert res["rsquared"] == pytest.approx(float(g["ml_rsquared"]), rel=1e-6)
    # K > N raises like the reference
    with pytest.raises(IndexError):
        mo.metrics_list(np.ones((5, 100, 3)), np.ones((5, 100, 3)))


def test_r2_matches_sklearn():
    from sklearn.metrics import r2_score
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((20, 7)), rng.standard_normal((20, 7))
    assert mo.r2_score_multi(a, b) == pytest.approx(r2_score(a, b), rel=1e-12)
    assert mo.r2_score_1d(a[:, 0], b[:, 0]) == pytest.approx(r2_score(a[:, 0], b[:, 0]), rel=1e-12)
    const = np.ones(10)
    assert mo.r2_score_1d(const, const) == r2_score(const, const) == 1.0
    assert mo.r2_score_1d(const, const + 1) == r2_score(const, const + 1) == 0.0


def test_zscore_onehot_and_frame_selection(golden_dir):
    g = _load(golden_dir, "metrics_kat.npz")
    mean, std = ro.zscore_stats(g["std_in"])
    np.testing.assert_array_equal(mean, g["std_mean"])
    np.testing.assert_array_equal(std, g["std_std"])
    assert std[2, 1] == 1e-8                                   # constant column clipped
    np.testing.assert_array_equal((g["std_in"] - mean) / std, g["std_z"])
    np.testing.assert_array_equal(ro.one_hot(g["oh_in"], 4), g["oh_out"])
    idx = ro.select_frames(42)
    np.testing.assert_array_equal(idx, g["sorted_idx"])        # bit-exact frame indexing
    sha = hashlib.sha256(idx.astype(np.int64).tobytes()).hexdigest()
    assert sha == "dd2daad4db659d6ba674faae1298feca07250cd19127edc5455eff606f2f36b3"
    assert 119 not in idx and len(idx) == 100 and np.all(np.diff(idx) > 0)


# ----------------------------------------------------------------------------- Linear
def test_onecycle_schedule(golden_dir):
    tab = _load(golden_dir, "onecycle.npz")["table"]
    for step, lr, b1 in tab:
        olr, ob1 = lo.one_cycle(int(step), 5000, 5e-5, 0.15, 10.0)
        assert olr == pytest.approx(lr, rel=1e-12, abs=1e-20)
        assert ob1 == pytest.approx(b1, rel=1e-12)
    assert tab[0][1] == pytest.approx(5e-6) and tab[0][2] == pytest.approx(0.95)


def test_linear_init_and_steps(golden_dir):
    g = _load(golden_dir, "linear_small.npz")
    w0 = _load(golden_dir, "linear_small_w0.npz")
    H, W, N = int(g["H"]), int(g["W"]), int(g["N"])
    params = lo.init_params(120 * H * W, N, seed=42)
    sd = lo.to_state_dict(params)
    # identical initial weights (same torch seed, same construction order)
    np.testing.assert_array_equal(sd["encoder.layers.0.weight"][:4, :64].numpy(), w0["w0_slice"])
    assert float(sd["encoder.layers.0.weight"].double().sum()) == pytest.approx(float(w0["w0_sum"]), rel=1e-12)
    for k in sd:
        if "init/" + k in g.files:
            np.testing.assert_array_equal(sd[k].numpy(), g["init/" + k])
    frames, ap = torch.from_numpy(g["frames"]), torch.from_numpy(g["ap"])
    logits = lo.forward(params, lo.cast_frames(frames[0])).reshape(-1, 100, N)
    np.testing.assert_allclose(logits.numpy(), g["first_logits"], rtol=1e-5, atol=1e-6)
    tr = lo.Trainer(params, total_steps=int(g["total_steps"]))
    for s in range(frames.shape[0]):
        lr, b1 = tr.hyper()
        assert lr == pytest.approx(float(g["lrs"][s]), rel=1e-12)
        assert b1 == pytest.approx(float(g["beta1s"][s]), rel=1e-12)
        loss = tr.step(frames[s], ap[s])
        assert loss == pytest.approx(float(g["losses"][s]), rel=2e-6)
    fsd = lo.to_state_dict(tr.params)
    for k in fsd:
        if "final/" + k in g.files:
            np.testing.assert_allclose(fsd[k].numpy(), g["final/" + k], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(fsd["encoder.layers.0.weight"].numpy()[::37, ::13], w0["w0_final_rows"], rtol=2e-4, atol=2e-7)
    rates = tr.predict_rates(frames[0], N)
    np.testing.assert_allclose(rates.numpy(), g["final_rates"], rtol=1e-4)


# ----------------------------------------------------------------------------- RRR
@pytest.mark.parametrize("name", ["a", "b"])
def test_rrr_against_reference(golden_dir, name):
    g = _load(golden_dir, "rrr_small.npz")
    sidx = g["sorted_idx"]
    data, gt = ro.preprocess_session([g[f"{name}/Xtr"], g[f"{name}/Xte"]], [g[f"{name}/ytr"], g[f"{name}/yte"]], sidx)
    np.testing.assert_allclose(data["X"][0][:3], g[f"{name}/X0_proc"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(data["y"][0][:3], g[f"{name}/y0_proc"], rtol=1e-12, atol=1e-12)
    td = {"e1": data}
    params, mse, trace = ro.train_model_main(td, 100.0, 3)
    assert len(trace) == 20                                    # 1 + 19 closure evaluations
    assert mse["mse_val_mean"] == pytest.approx(float(g[f"{name}/mse_val_mean"]), rel=1e-9)
    np.testing.assert_allclose(mse["mses_val"]["e1"], g[f"{name}/mses_val"], rtol=1e-8)
    np.testing.assert_allclose(params["e1_U"], g[f"{name}/U"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(params["V"], g[f"{name}/V"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(params["e1_b"], g[f"{name}/b"], rtol=1e-6, atol=1e-10)
    loss, _, _ = ro.loss_and_grad_dense(params, td, 100.0, 0)
    assert loss == pytest.approx(float(g[f"{name}/final_train_loss"]), rel=1e-9)
    _, _, pred = ro.predict_y_fr(params, td, "e1", 1)
    np.testing.assert_allclose(pred, g[f"{name}/pred_fr"], rtol=1e-7, atol=1e-9)
    # factorised closure == reference closure (value and every gradient)
    l1, g1, s1 = ro.loss_and_grad_dense(params, td, 100.0)
    l2, g2, s2 = ro.loss_and_grad_lowrank(params, td, 100.0)
    assert l2 == pytest.approx(l1, rel=1e-12)
    for k in g1:
        np.testing.assert_allclose(g2[k], g1[k], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(s2["e1"], s1["e1"], rtol=1e-12)
    # the autograd transcription of the reference closure (used as bench.py's CPU baseline) agrees too
    l3, g3 = ro.loss_and_grad_autograd(params, td, 100.0)
    assert l3 == pytest.approx(l1, rel=1e-12)
    for k in g1:
        np.testing.assert_allclose(g3[k], g1[k], rtol=1e-9, atol=1e-9)


def test_rrr_init_kat(golden_dir):
    g = _load(golden_dir, "rrr_small.npz")
    X = np.zeros((4, 100, 4)); y = np.zeros((4, 100, 2))
    p = ro.rrr_init({"s": {"X": [X], "y": [y]}}, 3)
    np.testing.assert_allclose(p["s_U"].ravel()[:4], g["init_U_first4"], rtol=0, atol=0)
    np.testing.assert_allclose(p["s_U"].ravel()[:4], [0.10184761, 0.02310309, 0.05650746, 0.12937803], atol=5e-9)
    assert list(p.keys()) == ["s_U", "s_b", "V"]
