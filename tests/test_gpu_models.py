"""-m gpu: model-level parity through the reference-shaped Python surface (which calls the C-ABI):
`Linear` train steps vs the oracle trainer and the committed reference outputs; RRR closure, fit and
prediction vs the float64 oracle and the committed reference outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import linear_oracle as lo
from oracle import rrr_oracle as ro
from tests.helpers import make_linear_model, small_rrr_problem

lo_synth = lo.synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vs(cuda):
    import vsb200
    return vsb200


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ----------------------------------------------------------------------------- Linear
def test_linear_state_dict_and_init_match_reference(vs, cuda, golden_dir):
    g = _load(golden_dir, "linear_small.npz"); w0 = _load(golden_dir, "linear_small_w0.npz")
    model, _, _ = make_linear_model(120 * 8 * 8, int(g["N"]), cuda)
    sd = model.state_dict()
    assert list(sd.keys()) == [f"{p}.layers.{i}.{w}" for p in ("encoder", "decoder") for i in (0, 2, 4) for w in ("weight", "bias")]
    np.testing.assert_array_equal(sd["encoder.layers.0.weight"][:4, :64].cpu().numpy(), w0["w0_slice"])
    for k in sd:
        if "init/" + k in g.files:
            np.testing.assert_array_equal(sd[k].cpu().numpy(), g["init/" + k])


@pytest.mark.parametrize("engine", [1, 2])
def test_linear_fused_steps_match_reference_golden(vs, cuda, golden_dir, engine):
    """6 optimizer steps on the reference's own recorded batches: per-step loss within rel 1e-3
    (BASELINE tolerance; observed ~1e-6), final weights and eval rates close."""
    g = _load(golden_dir, "linear_small.npz"); w0 = _load(golden_dir, "linear_small_w0.npz")
    N = int(g["N"])
    model, opt, sched = make_linear_model(120 * 8 * 8, N, cuda, total_steps=int(g["total_steps"]))
    model.engine = engine
    frames, ap = torch.from_numpy(g["frames"]).to(cuda), torch.from_numpy(g["ap"]).to(cuda)
    for s in range(frames.shape[0]):
        assert opt.param_groups[0]["lr"] == pytest.approx(float(g["lrs"][s]), rel=1e-12)
        assert opt.param_groups[0]["betas"][0] == pytest.approx(float(g["beta1s"][s]), rel=1e-12)
        loss = float(model.fused_train_step(frames[s], ap[s], opt))
        sched.step()
        assert loss == pytest.approx(float(g["losses"][s]), rel=1e-3)
        assert loss == pytest.approx(float(g["losses"][s]), rel=2e-5)      # what we actually achieve
    sd = model.state_dict()
    for k in sd:
        if "final/" + k in g.files:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g["final/" + k], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(sd["encoder.layers.0.weight"].cpu().numpy()[::37, ::13], w0["w0_final_rows"], rtol=2e-3, atol=2e-6)
    model.eval()
    with torch.no_grad():
        rates = torch.exp(model(frames[0]))
    assert rates.shape == (int(g["B"]), 100, N)
    np.testing.assert_allclose(rates.cpu().numpy(), g["final_rates"], rtol=1e-3)


def _assert_weights_close(oracle_params, model, steps=3, lr_max=5e-5):
    """Adam divides by sqrt(v): a weight whose gradient is ~0 moves by +-lr on rounding noise alone, so a
    handful of entries may differ by up to steps*lr; everything else must agree to float32 accuracy."""
    for (Wo, bo), (lin, _) in zip(oracle_params, model._layers):
        for got, ref in ((lin.weight.detach().cpu(), Wo), (lin.bias.detach().cpu(), bo)):
            diff = (got - ref).abs()
            assert float(diff.max()) <= 2.0 * steps * lr_max
            assert float((diff > 2e-6 + 2e-3 * ref.abs()).float().mean()) < 1e-3


def test_linear_autograd_route_with_stock_adamw(vs, cuda):
    """Reference-style loop (model -> criterion -> backward -> torch.optim.AdamW) on our model vs the oracle."""
    H = W = 12; N = 5; B = 6
    D = 120 * H * W
    model, opt, sched = make_linear_model(D, N, cuda, total_steps=30, fused=False)
    tr = lo.Trainer(lo.init_params(D, N, seed=42), total_steps=30)
    crit = torch.nn.PoissonNLLLoss(reduction="none", log_input=True)
    for s in range(3):
        frames, ap = lo.synth_batch(B, (120, 1, H, W), N, seed=10 + s)
        ref = tr.step(frames, ap)
        out = model(frames.to(cuda).float().flatten(1))           # float input like the reference loader
        loss = crit(out, ap.to(cuda)).mean()
        loss.backward()
        opt.step(); sched.step(); opt.zero_grad()
        assert float(loss) == pytest.approx(ref, rel=1e-4)
    _assert_weights_close(tr.params, model)


def test_linear_fused_optimizer_step_route(vs, cuda):
    """autograd backward + FusedAdamW.step(): the first-layer gradient stays factored (never materialised)."""
    H = W = 12; N = 5; B = 6
    D = 120 * H * W
    model, opt, sched = make_linear_model(D, N, cuda, total_steps=30, fused=True)
    tr = lo.Trainer(lo.init_params(D, N, seed=42), total_steps=30)
    crit = torch.nn.PoissonNLLLoss(reduction="none", log_input=True)
    for s in range(3):
        frames, ap = lo.synth_batch(B, (120, 1, H, W), N, seed=20 + s)
        ref = tr.step(frames, ap)
        out = model(frames.to(cuda))
        loss = crit(out, ap.to(cuda)).mean()
        loss.backward()
        w0 = model.encoder.layers[0].weight
        assert w0.grad is None and w0._vs_lowrank_grad is not None
        opt.step(); sched.step(); opt.zero_grad()
        assert float(loss) == pytest.approx(ref, rel=1e-4)
    _assert_weights_close(tr.params, model)


def test_linear_full_size_step_vs_oracle(vs, cuda):
    """BASELINE config 1 at full size (D = 1,966,080, N = 144, B = 16), sparse parity frames:
    2 steps, loss within rel 1e-3 of the fp32 oracle."""
    D, N, B = 120 * 128 * 128, 144, 16
    model, opt, sched = make_linear_model(D, N, cuda, total_steps=5000)
    tr = lo.Trainer(lo.init_params(D, N, seed=42), total_steps=5000)
    for s in range(2):
        frames, ap = lo.synth_batch(B, (120, 1, 128, 128), N, seed=s)
        ref = tr.step(frames, ap)
        got = float(model.fused_train_step(frames.to(cuda), ap.to(cuda), opt))
        sched.step()
        assert got == pytest.approx(ref, rel=1e-3)
    w_or = tr.params[0][0]
    w_gpu = model.encoder.layers[0].weight
    torch.testing.assert_close(w_gpu[:, ::4099].cpu(), w_or[:, ::4099], rtol=1e-3, atol=1e-6)


def test_model_pickles_like_reference_checkpoints(vs, cuda, tmp_path):
    model, _, _ = make_linear_model(120 * 8 * 8, 4, cuda)
    frames, _ = lo.synth_batch(2, (120, 1, 8, 8), 4)
    with torch.no_grad():
        a = model(frames.to(cuda))
    torch.save({"model": model, "epoch": 0}, tmp_path / "model_best.pt")           # src/trainer/base.py:285-291
    m2 = torch.load(tmp_path / "model_best.pt", weights_only=False)["model"]       # base.py:212
    with torch.no_grad():
        b = m2(frames.to(cuda))
    assert torch.equal(a, b)


# ----------------------------------------------------------------------------- RRR
def _params_to_model(model, params, dev):
    with torch.no_grad():
        for k, v in params.items():
            model.model[k].copy_(torch.from_numpy(v).to(dev))


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("planes,operand,rtol", [(1, "bf16", 2e-3), (1, "f16", 3e-4), (2, "bf16", 2e-5), (3, "bf16", 2e-6), (2, "f16", 2e-5)])
def test_rrr_closure_matches_oracle(vs, cuda, engine, planes, operand, rtol):
    """One closure evaluation (loss, per-neuron SSE, dU, dV, db) at perturbed parameters vs float64, for every operand
    mode: plain bf16 (8 bits), plain IEEE half (11 bits: the default fast mode), residual planes (parity modes)."""
    from model.rrr import RRRGD
    td = small_rrr_problem(seed=1, K=30, Kt=9, F=200, N=21)
    params = ro.rrr_init(td, 3)
    rng = np.random.default_rng(0)
    for k in params:
        params[k] = params[k] + 0.05 * rng.standard_normal(params[k].shape)
    loss_o, g_o, sse_o = ro.loss_and_grad_dense(params, td, 100.0, 0)
    m = RRRGD(td, 3, l2=100.0, planes=planes, engine=engine, operand=operand)
    m.to(cuda)
    _params_to_model(m, params, cuda)
    loss = m.loss_and_grad(td, 0)
    assert float(loss) == pytest.approx(loss_o, rel=max(rtol * 0.1, 1e-7))
    for k in g_o:
        got = m.model[k].grad.cpu().numpy()
        scale = np.abs(g_o[k]).max()
        assert np.abs(got - g_o[k]).max() <= rtol * scale, k
    sse = m.compute_MSE_RRRGD(td, 0)["e1"].cpu().numpy()
    np.testing.assert_allclose(sse, sse_o["e1"], rtol=max(rtol, 1e-6))
    # evaluation split + prediction
    _, _, yhat = m.predict_y(td, "e1", 1)
    ref = ro.predict(ro.compute_beta(params["e1_U"], params["V"], params["e1_b"]), td["e1"]["X"][1])
    err = yhat.cpu().numpy() - ref
    # plain bf16 operands: each product carries ~2^-9 relative rounding, so z-scored predictions agree to
    # ~2e-3 in relative L2 (loss to ~3e-4); plain half ~3e-4; two bf16 planes ~1e-5, three ~1e-6 (DESIGN.md "RRR precision")
    assert np.linalg.norm(err) <= 2.0 * rtol * np.linalg.norm(ref)
    assert np.abs(err).max() <= 3.0 * rtol * np.abs(ref).max()


@pytest.mark.parametrize("K,F,N", [(30, 200, 21), (37, 300, 144), (64, 129, 33), (16, 260, 160), (100, 520, 5), (130, 400, 150), (5, 300, 7),
                                   (70, 140, 64)])
def test_rrr_dense_backward_matches_factorised_and_oracle(vs, cuda, monkeypatch, K, F, N):
    """The per-time-bin dense backward (rrr_bwd_dense_kernel: D_t in TMEM, rank-one updates by the epilogue warps) against the
    factorised GEMM-B route and float64: trial counts that are no multiple of 16 (zero-padded residual operand, boxes that start
    at unaligned columns), odd numbers of 128-row feature tiles, every accumulator width up to 160 neurons."""
    from model.rrr import RRRGD
    td = small_rrr_problem(seed=K + N, K=K, Kt=5, F=F, N=N)
    params = ro.rrr_init(td, 3)
    rng = np.random.default_rng(0)
    for k in params:
        params[k] = params[k] + 0.05 * rng.standard_normal(params[k].shape)
    loss_o, g_o, _ = ro.loss_and_grad_dense(params, td, 100.0, 0)
    m = RRRGD(td, 3, l2=100.0, planes=1, engine=2); m.to(cuda)
    _params_to_model(m, params, cuda)
    got = {}
    for mode in ("0", "1", "2"):     # factorised GEMM-B / dense, half of the neurons per CTA / dense, CTA pairs (the default)
        monkeypatch.setenv("VS_RRR_DENSE", mode)
        loss = float(m.loss_and_grad(td, 0))
        assert loss == pytest.approx(loss_o, rel=2e-4)
        got[mode] = m.model["e1_U"].grad.cpu().numpy().copy()
    scale = np.abs(g_o["e1_U"]).max()
    for mode in got:
        assert np.abs(got[mode] - g_o["e1_U"]).max() <= 3e-3 * scale, mode
    assert np.abs(got["1"] - got["0"]).max() <= 5e-3 * scale        # two independent bf16 roundings (R vs R (x) V)
    assert not np.array_equal(got["1"], got["0"])          # the routes really are different kernels
    # both dense kernels contract the same bf16 operands bin by bin; they differ only in fp32 summation details
    assert np.abs(got["2"] - got["1"]).max() <= 1e-5 * scale


def test_rrr_closure_is_deterministic(vs, cuda):
    from model.rrr import RRRGD
    td = small_rrr_problem(seed=2, K=20, Kt=6, F=130, N=9)
    m = RRRGD(td, 3, l2=100.0, planes=1); m.to(cuda)
    a = m.loss_and_grad(td, 0); ga = m.model["e1_U"].grad.clone(); va = m.model["V"].grad.clone()
    b = m.loss_and_grad(td, 0)
    assert float(a) == float(b) and torch.equal(ga, m.model["e1_U"].grad) and torch.equal(va, m.model["V"].grad)


@pytest.mark.parametrize("name", ["a", "b"])
def test_rrr_fit_matches_reference_golden(vs, cuda, golden_dir, name):
    """Whole fit (init, one LBFGS.step, val SSE, de-z-scored prediction) vs the reference's recorded
    outputs, 3 operand planes: rel 1e-3 on loss and predictions (BASELINE tolerance)."""
    from model.rrr import train_model_main
    g = _load(golden_dir, "rrr_small.npz")
    data, gt = ro.preprocess_session([g[f"{name}/Xtr"], g[f"{name}/Xte"]], [g[f"{name}/ytr"], g[f"{name}/yte"]], g["sorted_idx"])
    td = {"e1": data}
    model, mse = train_model_main(td, l2=100, n_comp=3, model_fname="tmp", save=False, planes=3)
    assert model.n_closure_evals == 20
    assert float(mse["mse_val_mean"]) == pytest.approx(float(g[f"{name}/mse_val_mean"]), rel=1e-3)
    np.testing.assert_allclose(mse["mses_val"]["e1"].cpu().numpy(), g[f"{name}/mses_val"], rtol=1e-3)
    _, _, pred = model.predict_y_fr(td, "e1", 1)
    ref = g[f"{name}/pred_fr"]
    assert np.abs(pred.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    ev_o = ro.eval_session(ref, gt)
    ev = ro.eval_session(pred.cpu().numpy(), gt)
    assert ev["co_bps"] == pytest.approx(ev_o["co_bps"], rel=1e-3, abs=1e-5)
    assert ev["r2"] == pytest.approx(ev_o["r2"], rel=1e-3, abs=1e-5)


def test_rrr_single_plane_fits_track_the_reference(vs, cuda):
    """The fast single-plane modes through a whole fit.  Every closure evaluation is within 2e-3 (bf16) / 3e-4 (half) of
    float64 at the same parameters (test above); 20 un-line-searched L-BFGS iterations then amplify any perturbation
    (DESIGN.md 'RRR precision'), so the end of the fit is compared loosely here and tightly in the 3-plane golden test."""
    from model.rrr import train_model_main
    td = small_rrr_problem(seed=3, K=40, Kt=12, F=300, N=16)
    _, mse_o, _ = ro.train_model_main(td, 100.0, 3)
    for operand in ("bf16", "f16"):
        model, mse = train_model_main(td, l2=100.0, n_comp=3, model_fname="tmp", save=False, planes=1, operand=operand)
        assert model.fmt == (vs.OPERAND_F16 if operand == "f16" else vs.OPERAND_BF16)
        assert float(mse["mse_val_mean"]) == pytest.approx(mse_o["mse_val_mean"], rel=5e-2)


def test_rrr_half_operands_report_overflow(vs, cuda):
    """A feature that is constant in the train split is z-scored with std 1e-8: its test values explode
    past the half range.  The half mode must fail loudly; bf16 carries the reference's huge-but-finite numbers."""
    from model.rrr import RRRGD
    td = small_rrr_problem(seed=9, K=12, Kt=4, F=40, N=4)
    Xte = td["e1"]["X"][1]
    Xte[:, :, 3] = 2.5e8                                                        # what (x - mean) / 1e-8 produces
    m = RRRGD(td, 3, l2=100.0, planes=1, operand="f16"); m.to(cuda)
    m.loss_and_grad(td, 0)
    with pytest.raises(vs.VsError, match="half range"):
        m.compute_MSE_RRRGD(td, 1)
    b = RRRGD(td, 3, l2=100.0, planes=1, operand="bf16"); b.to(cuda)
    b.loss_and_grad(td, 0)
    assert torch.isfinite(b.compute_MSE_RRRGD(td, 1)["e1"]).all()


def test_rrr_multi_session_shared_V(vs, cuda):
    """Joint model over two sessions (V shared, rrr.py:46-49): dV accumulates over sessions."""
    from model.rrr import RRRGD
    td = {**small_rrr_problem(seed=4, K=16, Kt=5, F=70, N=6, eid="s1"), **small_rrr_problem(seed=5, K=12, Kt=5, F=90, N=11, eid="s2")}
    params = ro.rrr_init(td, 3)
    loss_o, g_o, _ = ro.loss_and_grad_dense(params, td, 100.0, 0)
    m = RRRGD(td, 3, l2=100.0, planes=3); m.to(cuda)
    for k, v in params.items():
        np.testing.assert_array_equal(m.model[k].detach().cpu().numpy(), v)          # same init stream
    loss = m.loss_and_grad(td, 0)
    assert float(loss) == pytest.approx(loss_o, rel=1e-6)
    gmax = max(np.abs(v).max() for v in g_o.values())
    for k in g_o:   # db is ~0 at the initial b = mean(y): compare on the scale of the whole gradient
        assert np.abs(m.model[k].grad.cpu().numpy() - g_o[k]).max() <= 1e-5 * np.abs(g_o[k]).max() + 1e-7 * gmax


def test_rrr_device_preprocessing_matches_oracle(vs, cuda):
    """R0 on the device from uint8 frames (colstats, frame gather, z-score, planes) and y smoothing."""
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=6, K=18, Kt=7, F=77, N=5, raw=True)
    data, _ = ro.preprocess_session([Xtr, Xte], [ytr, yte], sidx)
    K, Tf, F = Xtr.shape
    T = 100
    fr = torch.from_numpy(Xtr).to(cuda)
    mean = torch.empty(Tf * F, dtype=torch.float64, device=cuda); sd = torch.empty_like(mean)
    vs.check(vs.lib.vs_rrr_colstats(vs.ptr(fr), K, Tf * F, vs.ptr(mean), vs.ptr(sd), vs.stream()))
    np.testing.assert_allclose(mean.cpu().numpy().reshape(Tf, F), data["setup"]["mean_X_Tv"], rtol=1e-14)
    np.testing.assert_allclose(sd.cpu().numpy().reshape(Tf, F), data["setup"]["std_X_Tv"], rtol=1e-12)
    d = vs.RrrDims(K, T, F, 5, 3, 3, vs.lib.vs_rrr_ldc(F), vs.lib.vs_rrr_ldr(K, T))
    Xa = torch.zeros((3, K * T, d.ldc), dtype=torch.bfloat16, device=cuda)
    Xb = torch.full((3, F, d.ldr), 7.0, dtype=torch.bfloat16, device=cuda)      # the pack call must zero the pad itself
    xl = torch.empty(K * T, dtype=torch.float32, device=cuda)
    idx = torch.from_numpy(sidx.astype(np.int32)).to(cuda)
    vs.check(vs.lib.vs_rrr_pack_u8(vs.ptr(fr), Tf, vs.ptr(idx), vs.ptr(mean), vs.ptr(sd), d, vs.ptr(Xa), vs.ptr(Xb), vs.ptr(xl), None, vs.stream()))
    # operand rows are TIME-MAJOR (row d = t*K + k, include/vs_b200.h "RRR"): reorder the reference the same way
    ref = np.ascontiguousarray(data["X"][0][:, :, :-1].transpose(1, 0, 2)).reshape(T * K, F)
    got = Xa.double().sum(0)[:, :F].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-7)              # 3 bf16 planes ~ 24 bits
    # Xb: row c, column t*Kp + k with every time bin zero-padded to Kp = K rounded up to 16 trials
    Kp = (K + 15) // 16 * 16
    xb = Xb.double().sum(0)[:, :T * Kp].reshape(F, T, Kp).cpu().numpy()
    np.testing.assert_array_equal(xb[:, :, :K].reshape(F, T * K).T, got)
    assert not xb[:, :, K:].any()
    assert torch.all(xl == 1.0)
    # the first plane of the expansion is exactly bf16(X) -- frame gather is index-exact
    np.testing.assert_array_equal(Xa[0, :, :F].float().cpu().numpy(), torch.from_numpy(ref).to(torch.bfloat16).float().numpy())
    # the single-plane fast path does the z-score in fp32 ((x - mean) * (1/std)): same values up to one bf16 ulp on the
    # rare elements that sit on a rounding boundary
    d1 = vs.RrrDims(K, T, F, 5, 3, 1, vs.lib.vs_rrr_ldc(F), vs.lib.vs_rrr_ldr(K, T))
    Xa1 = torch.zeros((1, K * T, d1.ldc), dtype=torch.bfloat16, device=cuda)
    Xb1 = torch.zeros((1, F, d1.ldr), dtype=torch.bfloat16, device=cuda)
    vs.check(vs.lib.vs_rrr_pack_u8(vs.ptr(fr), Tf, vs.ptr(idx), vs.ptr(mean), vs.ptr(sd), d1, vs.ptr(Xa1), vs.ptr(Xb1), vs.ptr(xl), None, vs.stream()))
    want = torch.from_numpy(ref).to(torch.bfloat16).float().numpy()
    got1 = Xa1[0, :, :F].float().cpu().numpy()
    assert np.mean(got1 != want) < 2e-3 and np.abs(got1 - want).max() <= 2.0 ** -7 * np.abs(want).max()
    np.testing.assert_array_equal(Xb1[0, :, :T * Kp].float().reshape(F, T, Kp)[:, :, :K].reshape(F, T * K).cpu().numpy().T, got1)
    # y: gaussian smoothing + z-score
    cnt = torch.from_numpy(ytr).float().to(cuda)
    sm = torch.empty_like(cnt)
    vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, 5, 2.0, None, None, vs.ptr(sm), vs.stream()))
    my = torch.empty(T * 5, dtype=torch.float64, device=cuda); sy = torch.empty_like(my)
    vs.check(vs.lib.vs_colstats_f32(vs.ptr(sm), K, T * 5, vs.ptr(my), vs.ptr(sy), vs.stream()))
    yz = torch.empty_like(cnt)
    vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, 5, 2.0, vs.ptr(my), vs.ptr(sy), vs.ptr(yz), vs.stream()))
    np.testing.assert_allclose(my.cpu().numpy().reshape(T, 5), data["setup"]["mean_y_TN"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(yz.cpu().numpy(), data["y"][0], rtol=2e-4, atol=2e-5)


def test_selective_frame_upload_is_byte_exact_and_changes_nothing(vs, cuda):
    """vs_h2d_select_frames (only the T selected frames of a trial cross PCIe) vs uploading everything: the same bytes on the
    device, and pack_session_from_frames builds bit-identical operands either way (the statistics of a frame depend on that
    frame alone, src/train_rrr.py:108-171)."""
    from model.rrr import pack_session_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=9, K=21, Kt=6, F=131, N=7, raw=True)
    ftr, fte = torch.from_numpy(Xtr).pin_memory(), torch.from_numpy(Xte).pin_memory()
    idx = np.ascontiguousarray(sidx, dtype=np.int32)
    out = torch.zeros((ftr.shape[0], len(idx), ftr.shape[2]), dtype=torch.uint8, device=cuda)
    vs.check(vs.lib.vs_h2d_select_frames(ftr.data_ptr(), ftr.shape[0], ftr.shape[1], ftr.shape[2], idx.ctypes.data, len(idx), vs.ptr(out), vs.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ftr[:, torch.from_numpy(sidx.astype(np.int64))])
    bad = idx[::-1].copy()
    assert vs.lib.vs_h2d_select_frames(ftr.data_ptr(), ftr.shape[0], ftr.shape[1], ftr.shape[2], bad.ctypes.data, len(bad), vs.ptr(out), vs.stream()) != 0
    ctr, cte = torch.from_numpy(ytr).float(), torch.from_numpy(yte).float()
    a = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=1)                       # host frames: selective upload
    b = pack_session_from_frames(ftr.to(cuda), ctr, fte.to(cuda), cte, sidx, 3, planes=1)     # device frames: everything is there
    c = pack_session_from_frames(ftr, ctr, fte.to(cuda), cte, sidx, 3, planes=1)              # mixed
    torch.cuda.synchronize()
    for k in (0, 1):
        for other in (b, c):
            a["X"][k].wait_ready(); other["X"][k].wait_ready()
            K, T = a["X"][k].K, a["X"][k].T
            Kp = (K + 15) // 16 * 16
            F = ftr.shape[2]
            assert torch.equal(a["X"][k].Xa[:, :, :F], other["X"][k].Xa[:, :, :F])          # (columns F..ldc are padding, never read)
            assert torch.equal(a["X"][k].Xb[:, :, :T * Kp], other["X"][k].Xb[:, :, :T * Kp])
            assert torch.equal(a["X"][k].y, other["X"][k].y) and torch.equal(a["X"][k].xl, other["X"][k].xl)
    sel = torch.from_numpy(sidx.astype(np.int64)).to(cuda)
    ma = a["setup"]["mean_X_Tv"].view(ftr.shape[1], -1); mb = b["setup"]["mean_X_Tv"].view(ftr.shape[1], -1)
    assert torch.equal(ma[sel], mb[sel]) and torch.isnan(ma).sum() == (ftr.shape[1] - len(idx)) * ftr.shape[2]


def test_joint_model_through_sharded_optimizer_world1(vs, cuda):
    """parallel.train_joint_model on one rank == train_model with FusedLBFGS on the same joint two-session model
    (the world_size-2 equivalence of ShardedLBFGS is covered on CPU/gloo in tests/test_distributed_cpu.py and on two
    GPUs by tools/joint_2gpu_check.py)."""
    from model.rrr import RRRGD, train_model
    from parallel import train_joint_model
    td = {**small_rrr_problem(seed=4, K=16, Kt=5, F=70, N=6, eid="s1"), **small_rrr_problem(seed=5, K=12, Kt=5, F=90, N=11, eid="s2")}
    a = RRRGD(td, 3, l2=100.0, planes=3); a.to(cuda)
    _, ra = train_model(a, td, a.make_optimizer(), "tmp", save=False)
    b = RRRGD(td, 3, l2=100.0, planes=3); b.to(cuda)
    _, rb = train_joint_model(b, td)
    # (a) is device-driven, (b) host-driven: same decisions, fused-multiply-add contraction differs, and the
    # un-line-searched fit amplifies that rounding noise by ~1e2..1e3
    assert float(rb["mse_val_mean"]) == pytest.approx(float(ra["mse_val_mean"]), rel=1e-7)
    for k in a.model.keys():
        np.testing.assert_allclose(b.model[k].detach().cpu().numpy(), a.model[k].detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
    # a rank that holds only session s2 still draws s1's init from the stream (init_plan): same U_s2 and V
    plan = [(e, td[e]["y"][0].shape[2], td[e]["X"][0].shape[2], td[e]["y"][0].shape[1]) for e in td]
    c = RRRGD({"s2": td["s2"]}, 3, l2=100.0, planes=3, init_plan=plan)
    ref = RRRGD(td, 3, l2=100.0, planes=3)
    assert torch.equal(c.model["s2_U"], ref.model["s2_U"]) and torch.equal(c.model["V"], ref.model["V"])


# ----------------------------------------------------------------------------- RRR at BASELINE sizes
def _full_size_session(K, Kt, seed=0):
    import bench
    F, N = 110 * 166, 144
    ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, seed, pinned=False)
    return ftr, ctr, fte, cte, bench.sorted_idx_42()


def test_rrr_full_width_closure_vs_oracle(vs, cuda):
    """Full feature width (C = 110*166 + 1 = 18,261) and N = 144 neurons on a short session (K = 24), device R0 from
    uint8 frames, against the float64 oracle fed the same raw arrays: loss and every gradient."""
    from model.rrr import RRRGD, pack_session_from_frames
    ftr, ctr, fte, cte, sidx = _full_size_session(24, 8)
    data, _ = ro.preprocess_session([ftr.numpy(), fte.numpy()], [ctr.numpy().astype(np.float64), cte.numpy().astype(np.float64)], sidx)
    td_o = {"s": data}
    params = ro.rrr_init(td_o, 3)
    rng = np.random.default_rng(1)
    for k in params:
        params[k] = params[k] + 0.02 * rng.standard_normal(params[k].shape)
    loss_o, g_o, sse_o = ro.loss_and_grad_lowrank(params, td_o, 100.0, 0)
    # one bf16 plane: every residual carries the ~2^-8 rounding of its 18,260-term contraction and db averages only the
    # K = 24 trials of this short session, hence 1e-2 on the gradients here (2e-3 on the larger problems above)
    for planes, rtol in ((1, 1e-2), (3, 1e-5)):
        entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=planes, device=cuda)
        td = {"s": entry}
        m = RRRGD(td, 3, l2=100.0, planes=planes); m.to(cuda)
        _params_to_model(m, params, cuda)
        loss = float(m.loss_and_grad(td, 0))
        assert loss == pytest.approx(loss_o, rel=min(rtol, 1e-3))
        for k in g_o:
            got = m.model[k].grad.cpu().numpy()
            assert np.abs(got - g_o[k]).max() <= rtol * np.abs(g_o[k]).max(), (planes, k)
        del m, td, entry


def test_rrr_full_size_properties(vs, cuda):
    """BASELINE configs[1] at full size (K = 400, C = 18,261, N = 144, bf16 operands): properties that need no oracle.
    (a) bit-reproducible; (b) loss == sum of the per-neuron SSE + the L2 term evaluated independently with torch;
    (c) the gradient is the derivative of the loss: central difference along a random direction;
    (d) predict_y reproduces the closure's residuals."""
    from model.rrr import RRRGD, pack_session_from_frames
    ftr, ctr, fte, cte, sidx = _full_size_session(400, 80)
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=1, device=cuda)
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0, planes=1); m.to(cuda)
    g = torch.Generator(device="cpu").manual_seed(0)
    with torch.no_grad():
        for p in m.model.values():
            p.add_(0.02 * torch.randn(p.shape, generator=g, dtype=torch.float64).to(cuda))
    l0 = m.loss_and_grad(td, 0)
    grads = {k: p.grad.clone() for k, p in m.model.items()}
    l1 = m.loss_and_grad(td, 0)
    assert float(l0) == float(l1) and all(torch.equal(grads[k], m.model[k].grad) for k in grads)           # (a)
    sse = m.compute_MSE_RRRGD(td, 0)["s"]
    reg = m.regression_loss()["s"]
    assert float(l0) == pytest.approx(float(sse.sum() + reg), rel=1e-9)                                       # (b)
    _, y, yhat = m.predict_y(td, "s", 0)
    assert float(((yhat - y) ** 2).sum()) == pytest.approx(float(sse.sum()), rel=1e-5)                        # (d)
    # (c) with 3 operand planes (bf16 rounding of U makes the 1-plane loss piecewise; its gradient is checked against the
    # oracle at full width above): central difference along a random direction
    x0 = {k: p.detach().clone() for k, p in m.model.items()}
    del m, td, entry
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, planes=3, device=cuda)
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0, planes=3); m.to(cuda)
    with torch.no_grad():
        for k, p in m.model.items():
            p.copy_(x0[k])
    m.loss_and_grad(td, 0)
    grads = {k: p.grad.clone() for k, p in m.model.items()}
    d = {k: torch.randn(p.shape, generator=g, dtype=torch.float64).to(cuda) for k, p in m.model.items()}
    gd = sum(float((grads[k] * d[k]).sum()) for k in d)
    def central(eps):
        vals = []
        for sgn in (+1, -1):
            with torch.no_grad():
                for k, p in m.model.items():
                    p.copy_(x0[k] + sgn * eps * d[k])
            vals.append(float(m.loss_and_grad(td, 0)))
        return (vals[0] - vals[1]) / (2 * eps)
    # the loss is quartic in the parameters: Richardson extrapolation removes the O(eps^2) term of the central difference
    fd = (4.0 * central(1e-3) - central(2e-3)) / 3.0
    assert fd == pytest.approx(gd, rel=1e-4)


def test_linear_row_parallel_first_layer_emulated_ranks(vs, cuda):
    """Row-parallel first layer with 3 emulated ranks on one GPU: the partial pre-activations are summed by hand
    (what the NCCL all-reduce does), then every rank finishes the step locally.  The shards must reassemble the unsharded
    model: W0 slices, replicated layers identical on every rank, loss equal to the one-GPU fused step."""
    H = W = 20; N = 7; B = 16; world = 3
    D = 120 * H * W                                           # 48,000 pixels -> slices of 16,384 / 16,384 / 15,232
    ref, ropt, rsched = make_linear_model(D, N, cuda, total_steps=40)
    ranks = []
    for r in range(world):
        m, _, _ = make_linear_model(D, N, cuda, total_steps=40)           # same seed -> same full init on every rank
        lo, hi = m.shard_first_layer(r, world)
        from optim import FusedAdamW
        o = FusedAdamW(m.parameters(), lr=ropt.defaults["lr"], weight_decay=ropt.defaults["weight_decay"], eps=ropt.defaults["eps"])
        s = torch.optim.lr_scheduler.OneCycleLR(optimizer=o, total_steps=40, max_lr=5e-5, pct_start=0.15, div_factor=10)
        ranks.append((m, o, s, lo, hi))
    assert [(lo, hi) for *_, lo, hi in ranks] == [(0, 16384), (16384, 32768), (32768, 48000)]
    for step in range(3):
        frames, ap = lo_synth(B, (120, 1, H, W), N, seed=30 + step)
        frames, ap = frames.to(cuda), ap.to(cuda)
        want = float(ref.fused_train_step(frames, ap, ropt)); rsched.step()
        # phase 0 on every rank, emulated all-reduce, phase 1 on every rank
        # run the two phases by hand so the sum can be formed across the emulated ranks
        import ctypes as C
        ctx = []
        for m, o, s, lo, hi in ranks:
            x = frames.flatten(1)[:, lo:hi].contiguous()
            bufs = m.__dict__.setdefault("_train_bufs", {}).get((B, str(cuda)))
            if bufs is None:
                from model.linear import _MlpBuffers
                bufs = m.__dict__["_train_bufs"][(B, str(cuda))] = _MlpBuffers(m._layers, B, cuda, True)
            hyper = o.begin_fused_step()
            net = m._net(bufs, o.state)
            ws = m._workspace(net, B, cuda)
            args = (C.byref(net), vs.ptr(x), None, vs.ptr(ap), B, hyper, vs.ptr(bufs.loss), m.engine, vs.ptr(ws), ws.numel(), vs.stream())
            vs.check(vs.lib.vs_mlp_train_step_rowpar(*args, 0))
            ctx.append((args, bufs, net, x, ws))
        total = sum(c[1].act[0] for c in ctx)
        losses = []
        for (args, bufs, net, x, ws), (m, o, s, lo, hi) in zip(ctx, ranks):
            bufs.act[0].copy_(total)
            vs.check(vs.lib.vs_mlp_train_step_rowpar(*args, 1))
            s.step()
            losses.append(float(bufs.loss[0] / ap.numel()))
        assert all(l == losses[0] for l in losses)                       # replicas never drift
        assert losses[0] == pytest.approx(want, rel=1e-6)
    w_full = ref.encoder.layers[0].weight.detach()
    for m, o, s, lo, hi in ranks:
        torch.testing.assert_close(m.encoder.layers[0].weight.detach(), w_full[:, lo:hi], rtol=1e-4, atol=1e-7)
        for (la, _), (lb, _) in zip(m._layers[1:], ref._layers[1:]):
            torch.testing.assert_close(la.weight.detach(), lb.weight.detach(), rtol=1e-4, atol=1e-7)
            torch.testing.assert_close(la.bias.detach(), lb.bias.detach(), rtol=1e-4, atol=1e-7)
    # and the packaged call (world = 1 shard == the plain fused step)
    a, ao, _ = make_linear_model(D, N, cuda, total_steps=40)
    b, bo, _ = make_linear_model(D, N, cuda, total_steps=40)
    b.shard_first_layer(0, 1)
    frames, ap = lo_synth(B, (120, 1, H, W), N, seed=77)
    la = float(a.fused_train_step(frames.to(cuda), ap.to(cuda), ao))
    lb = float(b.fused_train_step_rowpar(frames.to(cuda), ap.to(cuda), bo))
    assert lb == pytest.approx(la, rel=1e-7)


@pytest.mark.parametrize("K,Kt,F,N,r", [(7, 3, 3, 5, 3), (33, 9, 1, 17, 2), (5, 2, 130, 1, 3), (19, 4, 65, 33, 5)])
def test_rrr_ragged_shapes_vs_oracle(vs, cuda, K, Kt, F, N, r):
    """Edge shapes: scalar modalities (C = 2..4 columns, src/train_rrr.py one-hot branch widths), odd trial counts,
    a single neuron, N and C1 that are no multiple of any tile size, rank != 3 -- closure and fit vs float64."""
    from model.rrr import RRRGD, train_model_main
    td = small_rrr_problem(seed=K + F, K=K, Kt=Kt, F=F, N=N)
    params = ro.rrr_init(td, r)
    rng = np.random.default_rng(2)
    for k in params:
        params[k] = params[k] + 0.05 * rng.standard_normal(params[k].shape)
    loss_o, g_o, sse_o = ro.loss_and_grad_dense(params, td, 100.0, 0)
    m = RRRGD(td, r, l2=100.0, planes=3); m.to(cuda)
    _params_to_model(m, params, cuda)
    assert float(m.loss_and_grad(td, 0)) == pytest.approx(loss_o, rel=1e-6)
    gmax = max(np.abs(v).max() for v in g_o.values())
    for k in g_o:
        assert np.abs(m.model[k].grad.cpu().numpy() - g_o[k]).max() <= 1e-5 * np.abs(g_o[k]).max() + 1e-7 * gmax, k
    _, mse_o, _ = ro.train_model_main(td, 100.0, r)
    _, mse = train_model_main(td, l2=100.0, n_comp=r, model_fname="tmp", save=False, planes=3)
    assert float(mse["mse_val_mean"]) == pytest.approx(mse_o["mse_val_mean"], rel=1e-3)


def test_linear_large_batch_step_vs_oracle(vs, cuda):
    """batch 48 (> 32: materialised-gradient route, tall-layer dW on the tensor cores) vs the fp32 oracle."""
    H = W = 12; N = 5; B = 48
    D = 120 * H * W
    model, opt, sched = make_linear_model(D, N, cuda, total_steps=30)
    tr = lo.Trainer(lo.init_params(D, N, seed=42), total_steps=30)
    for s in range(3):
        frames, ap = lo.synth_batch(B, (120, 1, H, W), N, seed=40 + s)
        ref = tr.step(frames, ap)
        got = float(model.fused_train_step(frames.to(cuda), ap.to(cuda), opt)); sched.step()
        assert got == pytest.approx(ref, rel=1e-4)
    _assert_weights_close(tr.params, model)


# ----------------------------------------------------------------------------- exact-operand mode (the default for uint8 frames)
@pytest.mark.parametrize("mode", ["exact", "dense"])
def test_rrr_exact_mode_closure_vs_oracle(vs, cuda, mode):
    """vs_rrr_pack_u8_exact + vs_rrr_closure_exact.  exact: z-score as hi+lo half planes forward (factorised GEMM), exact
    integer frames x hi+lo residual planes backward.  dense: both contractions on the exact integers, one time bin at a
    time (coefficient tiles generated on chip; dV from the second pass of the backward kernel).  float64 epilogues.
    One closure evaluation at full feature width against the float64 oracle: loss to 3e-5, every gradient to 1e-4 of
    its largest entry (one bf16 plane reaches 1e-2 on this problem).  What is left is the truncation of the fp32 TMEM
    accumulator over the ~3400 MMA steps of a 18,260-feature contraction: a smooth relative shrink of the prediction of
    ~1e-4, to which the whole fit is insensitive (profiles/r02_precision_sim_full.txt), unlike to operand rounding."""
    from model.rrr import RRRGD, pack_session_from_frames
    ftr, ctr, fte, cte, sidx = _full_size_session(24, 8)
    data, _ = ro.preprocess_session([ftr.numpy(), fte.numpy()], [ctr.numpy().astype(np.float64), cte.numpy().astype(np.float64)], sidx)
    td_o = {"s": data}
    params = ro.rrr_init(td_o, 3)
    rng = np.random.default_rng(1)
    for k in params:
        params[k] = params[k] + 0.02 * rng.standard_normal(params[k].shape)
    loss_o, g_o, sse_o = ro.loss_and_grad_lowrank(params, td_o, 100.0, 0)
    entry = pack_session_from_frames(ftr, ctr, fte, cte, sidx, 3, device=cuda, mode=mode)
    assert entry["X"][0].dims.mode == (vs.RRR_MODE_EXACT if mode == "exact" else vs.RRR_MODE_DENSE) and entry["X"][1].Xb is None
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0); m.to(cuda)
    assert m.exact and m.planes == 2
    _params_to_model(m, params, cuda)
    loss = float(m.loss_and_grad(td, 0))
    assert loss == pytest.approx(loss_o, rel=3e-5)
    for k in g_o:
        got = m.model[k].grad.cpu().numpy()
        assert np.abs(got - g_o[k]).max() <= 1e-4 * np.abs(g_o[k]).max(), k
    sse = m.compute_MSE_RRRGD(td, 0)["s"].cpu().numpy()
    np.testing.assert_allclose(sse, sse_o["s"], rtol=1e-4)
    # the validation split has no backward operand: evaluation works, a gradient request is refused loudly
    val = m.compute_MSE_RRRGD(td, 1)["s"]
    _, _, sse_val_o = ro.loss_and_grad_lowrank(params, td_o, 100.0, 1)
    np.testing.assert_allclose(val.cpu().numpy(), sse_val_o["s"], rtol=1e-4)
    with pytest.raises(vs.VsError):
        m.loss_and_grad(td, 1)
    # predict_y reproduces the closure's residuals
    _, yv, yhat = m.predict_y(td, "s", 0)
    assert float(((yhat - yv) ** 2).sum()) == pytest.approx(float(sse.sum()), rel=1e-6)


@pytest.mark.parametrize("mode", ["exact", "dense"])
@pytest.mark.parametrize("K,F,N", [(40, 160, 10), (37, 200, 33), (70, 130, 144), (300, 257, 16)])
def test_rrr_exact_mode_ragged_shapes(vs, cuda, K, F, N, mode):
    """Exact-operand mode on shapes that do not fill tiles (K not a multiple of 16, F not of 128, N not of 16)."""
    from model.rrr import RRRGD, pack_session_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=3, K=K, Kt=9, F=F, N=N, raw=True)
    data, _ = ro.preprocess_session([Xtr, Xte], [ytr, yte], sidx)
    td_o = {"s": data}
    params = ro.rrr_init(td_o, 3)
    rng = np.random.default_rng(2)
    for k in params:
        params[k] = params[k] + 0.05 * rng.standard_normal(params[k].shape)
    loss_o, g_o, _ = ro.loss_and_grad_lowrank(params, td_o, 100.0, 0)
    entry = pack_session_from_frames(torch.from_numpy(Xtr), torch.from_numpy(ytr), torch.from_numpy(Xte), torch.from_numpy(yte), sidx, 3,
                                     device=cuda, mode=mode)
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0); m.to(cuda)
    _params_to_model(m, params, cuda)
    # short contractions (F <= 257): no visible accumulator truncation, what is left is the 2^-22 of the two-plane operands
    assert float(m.loss_and_grad(td, 0)) == pytest.approx(loss_o, rel=2e-6)
    for k in g_o:
        got = m.model[k].grad.cpu().numpy()
        assert np.abs(got - g_o[k]).max() <= 2e-5 * np.abs(g_o[k]).max(), k


@pytest.mark.parametrize("mode", [None, "exact", "dense"])
def test_rrr_exact_mode_whole_fit_small_vs_oracle(vs, cuda, mode):
    """train_model_from_frames in its default mode and in both exact-operand modes against the oracle's float64 fit of the
    same raw arrays: validation SSE, de-z-scored predictions, co-bps and R2 (src/train_rrr.py:193-236)."""
    from model.rrr import train_model_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=0, K=40, Kt=12, F=160, N=10, raw=True)
    data, gt = ro.preprocess_session([Xtr, Xte], [ytr, yte], sidx)
    td_o = {"session": data}
    p_o, mse_o, _ = ro.train_model_main(td_o, 100.0, 3)
    model, mse, td = train_model_from_frames(torch.from_numpy(Xtr), torch.from_numpy(ytr), torch.from_numpy(Xte), torch.from_numpy(yte), sidx,
                                             l2=100.0, n_comp=3, mode=mode)
    assert model.exact
    assert float(mse["mse_val_mean"]) == pytest.approx(mse_o["mse_val_mean"], rel=1e-5)
    _, _, pred = model.predict_y_fr(td, "session", 1)
    _, _, pred_o = ro.predict_y_fr(p_o, td_o, "session", 1)
    # single predictions: the un-line-searched fit amplifies the last bits of a closure evaluation by 3-4 orders of magnitude
    # (DESIGN.md "RRR precision"), so elementwise agreement is looser than that of the sums BASELINE.json puts a tolerance on
    np.testing.assert_allclose(pred.cpu().numpy(), pred_o, rtol=2e-3, atol=1e-3)
    ev, ev_o = ro.eval_session(pred.cpu().numpy(), gt), ro.eval_session(pred_o, gt)
    assert ev["co_bps"] == pytest.approx(ev_o["co_bps"], rel=1e-3, abs=1e-6)
    assert ev["r2"] == pytest.approx(ev_o["r2"], rel=1e-3, abs=1e-6)


def test_rrr_default_mode_full_size_fit_within_tolerance(vs, cuda):
    """BASELINE configs[1] at FULL size (K = 400, C = 18,261, N = 144): the default mode of train_model_from_frames -- the
    one bench.py times -- against an independent float64 dense fit (torch einsum + autograd + torch.optim.LBFGS on the
    GPU, bench.fp64_dense_reference).  The reference trains with ONE un-line-searched LBFGS.step whose trajectory
    amplifies a 1e-6 perturbation of a closure evaluation to ~3e-5 of the result (profiles/r02_precision_sim_full.txt);
    tolerance rel 1e-3 on validation SSE, co-bps and R2 as BASELINE.json states."""
    import bench
    from model.rrr import train_model_from_frames
    ftr, ctr, fte, cte, sidx = _full_size_session(400, 80)
    ref = bench.fp64_dense_reference(ftr, ctr, fte, cte, sidx, cuda)
    model, mse, td = train_model_from_frames(ftr, ctr, fte, cte, sidx, l2=100.0, n_comp=3)
    assert model.exact
    got = float(mse["mse_val_mean"])
    assert abs(got - ref["val_sse"]) <= 1e-3 * ref["val_sse"], (got, ref["val_sse"])
    assert model.n_closure_evals >= 20
    # co-bps and per-trial R2 of the de-z-scored, clipped predictions against the unsmoothed test counts
    # (src/train_rrr.py:193-236); both sit near zero on this signal-free session, hence the absolute floors
    _, _, pred = model.predict_y_fr(td, "session", 1)
    gt = cte.numpy().astype(np.float64)
    ev, ev_o = ro.eval_session(pred.cpu().numpy(), gt), ro.eval_session(ref["pred_test_fr"], gt)
    assert ev["co_bps"] == pytest.approx(ev_o["co_bps"], rel=1e-3, abs=2e-5), (ev["co_bps"], ev_o["co_bps"])
    assert ev["r2"] == pytest.approx(ev_o["r2"], rel=1e-3, abs=2e-5), (ev["r2"], ev_o["r2"])


@pytest.mark.parametrize("mode", ["exact", "dense"])
@pytest.mark.parametrize("K,F", [(37, 200), (400, 18260 // 4)])
def test_fused_loader_matches_two_kernel_path(vs, cuda, monkeypatch, mode, K, F):
    """vs_rrr_pack_u8_fused (one read of the frames: statistics + pack per time bin, 128-bit loads and stores) against
    vs_rrr_colstats + vs_rrr_pack_u8_exact: statistics and every operand bit for bit."""
    from model.rrr import pack_session_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=7, K=K, Kt=11, F=F, N=16, raw=True)
    args = (torch.from_numpy(Xtr), torch.from_numpy(ytr), torch.from_numpy(Xte), torch.from_numpy(yte), sidx, 3)
    monkeypatch.setenv("VS_RRR_FUSED_PACK", "1")
    a = pack_session_from_frames(*args, device=cuda, mode=mode)
    monkeypatch.setenv("VS_RRR_FUSED_PACK", "0")
    b = pack_session_from_frames(*args, device=cuda, mode=mode)
    torch.cuda.synchronize()
    for k in ("mean_X_Tv", "std_X_Tv"):
        sel = torch.as_tensor(np.asarray(sidx), device=cuda).long()
        assert torch.equal(a["setup"][k].view(120, F)[sel], b["setup"][k].view(120, F)[sel]), k
    for which in (0, 1):
        sa, sb = a["X"][which], b["X"][which]
        sa.wait_ready(); sb.wait_ready()
        C1 = sa.C1
        if mode == "exact":
            assert torch.equal(sa.Xa[:, :, :C1].view(torch.int16), sb.Xa[:, :, :C1].view(torch.int16))
        else:
            assert torch.equal(sa.exact["Xc"][:, :C1].view(torch.int16), sb.exact["Xc"][:, :C1].view(torch.int16))
        if which == 0:
            Kp = (sa.K + 15) // 16 * 16
            assert torch.equal(sa.Xb[:, :sa.T * Kp].view(torch.int16), sb.Xb[:, :sa.T * Kp].view(torch.int16))
            for k in ("isdT", "qT"):
                assert torch.equal(sa.exact[k], sb.exact[k]), k
        assert torch.equal(sa.xl, sb.xl)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_loader_reports_train_constant_features_that_vary_elsewhere(vs, cuda, monkeypatch, fused):
    """a feature that is constant in the train split has its std clipped to 1e-8 (src/utils/utils.py:110); a
    test frame that differs there becomes ~1e8 after the z-score and leaves the IEEE-half range.  Both loaders (fused:
    decided once per column from the column extremes; two-kernel: per element) must raise the flag for the test split and
    for that split only; a feature that is constant EVERYWHERE is zero after the z-score and raises nothing."""
    from model.rrr import pack_session_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=3, K=24, Kt=9, F=160, N=8, raw=True)
    Xtr[:, :, 5] = 7; Xte[:, :, 5] = 7                      # constant in both splits: harmless
    Xtr[:, :, 131] = 200                                      # constant in train ...
    monkeypatch.setenv("VS_RRR_FUSED_PACK", fused)
    args = lambda te: (torch.from_numpy(Xtr), torch.from_numpy(ytr), torch.from_numpy(te), torch.from_numpy(yte), sidx, 3)
    Xte_ok = Xte.copy(); Xte_ok[:, :, 131] = 200
    ok = pack_session_from_frames(*args(Xte_ok), device=cuda, mode="exact")
    torch.cuda.synchronize()
    for sp in ok["X"]:
        sp.wait_ready()
        assert int(sp.overflow.item()) == 0
    assert torch.all(ok["X"][0].Xa[:, :, 5].view(torch.int16) == 0) and torch.all(ok["X"][1].Xa[:, :, 131].view(torch.int16) == 0)
    bad = pack_session_from_frames(*args(Xte), device=cuda, mode="exact")      # ... and varying in test
    torch.cuda.synchronize()
    bad["X"][1].wait_ready()
    assert int(bad["X"][0].overflow.item()) == 0
    assert int(bad["X"][1].overflow.item()) == 1
    with pytest.raises(vs.VsError):
        bad["X"][1].check_range()


def test_rrr_exact_mode_wide_session_runs_in_neuron_groups(vs, cuda):
    """More neurons than one launch of the exact-operand kernels holds (160; BASELINE lists sessions with N = 436): the
    default mode evaluates the session group by group -- the model is separable over neurons given V -- instead of falling
    back to the three-plane classic mode.  One closure evaluation (loss, every gradient, per-neuron SSE), predictions and
    the whole fit against the float64 oracle."""
    from model.rrr import RRRGD, pack_session_from_frames, train_model_from_frames
    Xtr, Xte, ytr, yte, sidx = small_rrr_problem(seed=11, K=40, Kt=12, F=160, N=170, raw=True)
    data, _ = ro.preprocess_session([Xtr, Xte], [ytr, yte], sidx)
    td_o = {"s": data}
    params = ro.rrr_init(td_o, 3)
    rng = np.random.default_rng(4)
    for k in params:
        params[k] = params[k] + 0.05 * rng.standard_normal(params[k].shape)
    loss_o, g_o, sse_o = ro.loss_and_grad_lowrank(params, td_o, 100.0, 0)
    args = (torch.from_numpy(Xtr), torch.from_numpy(ytr), torch.from_numpy(Xte), torch.from_numpy(yte), sidx)
    entry = pack_session_from_frames(*args, 3, device=cuda)
    sp = entry["X"][0]
    assert sp.dims.mode == vs.RRR_MODE_EXACT and [(a, b) for a, b, *_ in sp.neuron_groups()] == [(0, 96), (96, 170)]
    td = {"s": entry}
    m = RRRGD(td, 3, l2=100.0); m.to(cuda)
    assert m.exact
    _params_to_model(m, params, cuda)
    assert float(m.loss_and_grad(td, 0)) == pytest.approx(loss_o, rel=2e-6)
    for k in g_o:
        got = m.model[k].grad.cpu().numpy()
        assert np.abs(got - g_o[k]).max() <= 2e-5 * np.abs(g_o[k]).max(), k
    np.testing.assert_allclose(m.compute_MSE_RRRGD(td, 0)["s"].cpu().numpy(), sse_o["s"], rtol=1e-5)
    _, yv, yhat = m.predict_y(td, "s", 1)
    _, _, sse_val_o = ro.loss_and_grad_lowrank(params, td_o, 100.0, 1)
    np.testing.assert_allclose(((yhat - yv) ** 2).sum((0, 1)).cpu().numpy(), sse_val_o["s"], rtol=1e-5)
    # whole fit through the public entry point
    _, mse_o, _ = ro.train_model_main({"session": data}, 100.0, 3)
    model, mse, _ = train_model_from_frames(*args, l2=100.0, n_comp=3)
    assert model.exact
    assert float(mse["mse_val_mean"]) == pytest.approx(mse_o["mse_val_mean"], rel=1e-4)
