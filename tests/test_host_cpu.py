"""CPU: host-side logic and the C-ABI surface (no kernel is launched here: there is no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", hdr)))


def test_library_loads_and_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "video-spike_b200", "lib", "libvs_b200.so")
    assert os.path.exists(lib_path), "build it: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vs_b200.h but not exported"
    lib.vs_version.restype = ctypes.c_int
    hdr = open(os.path.join(ROOT, "include", "vs_b200.h")).read()
    assert lib.vs_version() == int(re.search(r"#define VS_ABI_VERSION (\d+)", hdr).group(1))


def test_python_binding_covers_the_header():
    import vsb200
    assert sorted(vsb200.EXPORTS) == _declared_symbols()


def test_ctypes_structs_mirror_the_header():
    """The device-resident L-BFGS state is parsed on the host: the ctypes mirror must have the C layout."""
    import ctypes as C
    import vsb200 as vs
    assert C.sizeof(vs.LbfgsDev) == vs.lib.vs_lbfgs_dev_state_bytes()
    host = vs.LbfgsDev()
    assert vs.lib.vs_lbfgs_dev_init_host(C.byref(host), 46) == 0
    assert (host.n_free, host.free_slots[45], host.free_slots[0], host.H_diag, host.m, host.done) == (46, 0, 45, 1.0, 0, 0)
    assert vs.LbfgsDev.dmax.offset + 8 <= 1728 < vs.LbfgsDev.out.offset      # what FusedLBFGS reads back / re-uploads


def test_no_cpu_fallback():
    """CPU tensors are refused by every wrapper: the product path cannot silently run without the GPU."""
    import vsb200 as vs
    with pytest.raises(vs.VsError):
        vs.ptr(torch.zeros(4))
    with pytest.raises(vs.VsError):
        vs.require_b200()
    from tests.helpers import make_linear_model
    model, opt, _ = make_linear_model(120 * 4 * 4, 2, torch.device("cpu"))
    with pytest.raises(vs.VsError):
        model(torch.zeros(2, 120 * 4 * 4))
    with pytest.raises(vs.VsError):
        model.fused_train_step(torch.zeros(2, 120 * 4 * 4, dtype=torch.uint8), torch.zeros(2, 100, 2), opt)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "video-spike_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src, os.path.join(dirpath, f)


# ----------------------------------------------------------------------------- L-BFGS coefficient-space recursion
def _explicit_two_loop(g, S, Y, H):
    """torch.optim.LBFGS.step's direction computation, on explicit vectors."""
    m = len(S)
    ro = [1.0 / float(Y[i] @ S[i]) for i in range(m)]
    al = [0.0] * m
    q = -g.copy()
    for i in range(m - 1, -1, -1):
        al[i] = float(S[i] @ q) * ro[i]
        q -= al[i] * Y[i]
    r = q * H
    for i in range(m):
        be = float(Y[i] @ r) * ro[i]
        r += (al[i] - be) * S[i]
    return r


@pytest.mark.parametrize("m", [0, 1, 4, 19])
def test_lbfgs_two_loop_matches_vector_recursion(m):
    from optim import lbfgs_two_loop
    rng = np.random.default_rng(m)
    n = 300
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n)      # SPD "Hessian": y = A s keeps y.s > 0
    S = [rng.standard_normal(n) for _ in range(m)]
    Y = [A @ s for s in S]
    g = rng.standard_normal(n)
    H = 0.37
    d_ref = _explicit_two_loop(g, S, Y, H)
    SY = [[float(S[i] @ Y[j]) for j in range(m)] for i in range(m)]
    YY = [[float(Y[i] @ Y[j]) for j in range(m)] for i in range(m)]
    cg, cs, cy, gtd = lbfgs_two_loop(float(g @ g), [float(s @ g) for s in S], [float(y @ g) for y in Y], SY, YY, H)
    d = cg * g
    for i in range(m):
        d = d + cs[i] * S[i] + cy[i] * Y[i]
    np.testing.assert_allclose(d, d_ref, rtol=1e-9, atol=1e-9 * np.abs(d_ref).max())
    assert gtd == pytest.approx(float(g @ d_ref), rel=1e-9)


@pytest.mark.parametrize("m", [0, 1, 7, 30])
def test_compiled_two_loop_equals_python_statement(m):
    import ctypes as C
    import vsb200 as vs
    from optim import lbfgs_two_loop
    rng = np.random.default_rng(m + 10)
    n, ld = 200, 40
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n)
    S = rng.standard_normal((m, n)); Y = S @ A
    g = rng.standard_normal(n)
    SY = np.zeros((ld, ld)); YY = np.zeros((ld, ld))
    SY[:m, :m] = S @ Y.T; YY[:m, :m] = Y @ Y.T
    sg, yg = S @ g, Y @ g
    cg, cs, cy, gtd = lbfgs_two_loop(float(g @ g), list(sg), list(yg), SY[:m, :m].tolist(), YY[:m, :m].tolist(), 0.7)
    coef = np.zeros(2 * m + 1); out = C.c_double()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)                                      # noqa: E731
    vs.check(vs.lib.vs_host_lbfgs_two_loop(m, float(g @ g), vp(np.ascontiguousarray(sg)), vp(np.ascontiguousarray(yg)), vp(SY), vp(YY), ld, 0.7,
                                           vp(coef), C.byref(out)))
    np.testing.assert_allclose(coef, [cg, *cs, *cy], rtol=1e-12, atol=1e-300)
    assert out.value == pytest.approx(gtd, rel=1e-12)


# ----------------------------------------------------------------------------- config / schedule host logic
def test_config_include_and_update_semantics():
    """src/utils/config_utils.py:20-42: `include:` pulls a YAML file in, update_config(path) merges, and
    update_config(args, config) is a no-op."""
    from tests.helpers import linear_config
    cfg = linear_config(1234, 7)
    assert cfg.model.encoder.input_dim == 1234 and cfg.model.decoder.output_dim == 700
    assert cfg.model.encoder.hidden_dims == [256, 128] and cfg.model.decoder.hidden_dims == [128, 256]
    assert cfg.optimizer.lr == pytest.approx(5e-5) and cfg.optimizer.wd == pytest.approx(0.01)
    assert cfg.training.train_batch_size == 16 and cfg.seed == 42


def test_model_structure_matches_reference_layout():
    from tests.helpers import make_linear_model
    model, opt, sched = make_linear_model(120 * 4 * 4, 3, torch.device("cpu"))
    keys = list(model.state_dict().keys())
    assert keys == [f"{p}.layers.{i}.{w}" for p in ("encoder", "decoder") for i in (0, 2, 4) for w in ("weight", "bias")]
    assert model.state_dict()["encoder.layers.0.weight"].shape == (256, 1920)
    assert model.state_dict()["decoder.layers.4.weight"].shape == (300, 256)
    assert model.output_dim == 3
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-6)            # OneCycleLR start = max_lr / div_factor
    assert opt.param_groups[0]["betas"][0] == pytest.approx(0.95)       # cycle_momentum drives beta1


# ----------------------------------------------------------------------------- RRR initialisation stream (R1)
@pytest.mark.parametrize("shapes", [[(7, 50, 3), (3, 100)], [(1,), (2,), (3,), (1, 1), (5000, 3)], [(3, 5, 3), (3, 100)] * 3])
def test_host_normal_stream_is_numpys_legacy_stream_bit_for_bit(shapes):
    """csrc/host_rng.cpp vs np.random.seed(0); np.random.normal(size)/sqrt(300) (src/model/rrr.py:35,42-43):
    identical doubles, identical generator state afterwards (key, pos, cached gaussian), any thread count."""
    import vsb200 as vs
    div = float(np.sqrt(300))
    np.random.seed(0)
    ref = [np.random.normal(size=s) / np.sqrt(300) for s in shapes]
    st_ref = np.random.get_state()
    for threads in (1, 3, 0):
        g = vs.LegacyNormalStream(0)
        got = [g.normal(s, div, threads) for s in shapes]
        for a, b in zip(ref, got):
            np.testing.assert_array_equal(a, b)
        np.random.seed(12345)
        g.export_to_numpy()
        st = np.random.get_state()
        assert st[0] == st_ref[0] and st[2:] == st_ref[2:]
        np.testing.assert_array_equal(st[1], st_ref[1])


def test_host_normal_stream_long_run_and_other_seed():
    import vsb200 as vs
    np.random.seed(42)
    ref = np.random.normal(size=1_000_003)
    tail = np.random.normal(size=5)
    g = vs.LegacyNormalStream(42)
    np.testing.assert_array_equal(g.normal((1_000_003,)), ref)
    np.testing.assert_array_equal(g.normal((5,)), tail)          # odd count: the cached second value comes first


def test_rrrgd_init_equals_reference_init():
    """RRRGD.__init__ (rrr.py:30-54) through the host generator == the oracle's numpy transcription, and the
    global numpy stream is left where the reference leaves it."""
    from model.rrr import RRRGD
    from oracle import rrr_oracle as ro
    from tests.helpers import small_rrr_problem
    td = {**small_rrr_problem(seed=4, K=16, Kt=5, F=70, N=6, eid="s1"), **small_rrr_problem(seed=5, K=12, Kt=5, F=90, N=11, eid="s2")}
    ref = ro.rrr_init(td, 3)
    after_ref = np.random.normal()
    m = RRRGD(td, 3, l2=100.0)
    after = np.random.normal()
    assert set(m.model.keys()) == set(ref.keys())       # nn.ParameterDict sorts a plain dict by key, here as in the reference
    for k, v in ref.items():
        np.testing.assert_array_equal(m.model[k].detach().numpy(), v)
    assert after == after_ref


# ----------------------------------------------------------------------------- frame windows (L0)
def test_load_video_index_matches_oracle_bit_exact():
    from oracle import loader_oracle as lo
    from utils.dataset_utils import load_video_index
    rng = np.random.default_rng(3)
    fps = 60.0
    ts = np.cumsum(rng.uniform(0.9, 1.1, size=20000) / fps)             # jittered camera clock
    t0 = np.sort(rng.uniform(ts[200], ts[-400], size=57))
    intervals = np.stack([t0, t0 + 2.0], axis=1)
    ref = lo.load_video_index(ts, intervals, fps)
    got = load_video_index(ts, intervals, fps)
    assert got.dtype == np.int64 and got.shape == (57, 120)
    np.testing.assert_array_equal(got, ref)
    # a start that coincides with a timestamp takes that very frame (searchsorted side='left')
    intervals[0] = (ts[1000], ts[1000] + 2.0)
    assert load_video_index(ts, intervals, fps)[0, 0] == 1000 == lo.load_video_index(ts, intervals, fps)[0, 0]
    # dropped frames -> the reference raises, so do we
    bad = np.delete(ts, np.arange(3000, 3030))
    with pytest.raises(ValueError):
        load_video_index(bad, np.array([[bad[2990], bad[2990] + 2.0]]), fps)
    with pytest.raises(ValueError):
        lo.load_video_index(bad, np.array([[bad[2990], bad[2990] + 2.0]]), fps)


def test_select_frames_is_the_reference_draw(golden_dir):
    from utils.utils import select_frames, set_seed
    g = np.load(os.path.join(golden_dir, "metrics_kat.npz"))
    set_seed(42)
    np.testing.assert_array_equal(select_frames(), g["sorted_idx"])


def test_rrr_pitch_helpers():
    """vs_rrr_ldc / vs_rrr_ldr are host-only: Xa rows are padded to 64 features, the backward operand pads every time bin to
    16 trials (a bin starts 32-byte aligned and ends on a tensor-core K-step) and its rows to 64 columns."""
    import vsb200 as vs
    for C1 in (1, 64, 65, 18260):
        assert vs.lib.vs_rrr_ldc(C1) == (C1 + 63) // 64 * 64
    for K, T in ((1, 1), (5, 100), (16, 100), (37, 100), (400, 100), (401, 7)):
        Kp = (K + 15) // 16 * 16
        assert vs.lib.vs_rrr_ldr(K, T) == (T * Kp + 63) // 64 * 64
        assert vs.lib.vs_rrr_ldr(K, T) >= T * K


def test_argument_validation_needs_no_gpu():
    """The C-ABI validates its arguments before it touches CUDA: empty / inconsistent shapes and null buffers come back as
    error codes with a message (never an exception, never a launch), which is also what a CPU-only box observes."""
    import vsb200 as vs
    C1, K, T, N = 200, 30, 100, 21
    good = vs.RrrDims(K, T, C1, N, 3, 1, vs.lib.vs_rrr_ldc(C1), vs.lib.vs_rrr_ldr(K, T), 0)
    assert vs.lib.vs_rrr_workspace(good) > 0
    for bad in (vs.RrrDims(0, T, C1, N, 3, 1, good.ldc, good.ldr, 0),          # no trials
                vs.RrrDims(K, T, C1, 0, 3, 1, good.ldc, good.ldr, 0)):         # no neurons
        assert vs.lib.vs_rrr_workspace(bad) == 0
    def closure_rc(d):
        return vs.lib.vs_rrr_closure(d, None, None, None, None, None, None, None, 100.0, None, None, None, None, None, 0, None, 0, None)
    assert closure_rc(vs.RrrDims(0, T, C1, N, 3, 1, good.ldc, good.ldr, 0)) != 0
    assert b"empty" in vs.lib.vs_last_error()
    assert closure_rc(vs.RrrDims(K, T, C1, N, 3, 4, good.ldc, good.ldr, 0)) != 0            # planes out of range
    assert closure_rc(vs.RrrDims(K, T, C1, N, 3, 1, good.ldc, K * T, 0)) != 0               # pitch ignores the padded time bins
    assert b"pitch" in vs.lib.vs_last_error()
    assert closure_rc(vs.RrrDims(K, T, C1, N, 3, 1, good.ldc, good.ldr, 7)) != 0            # unknown operand format
    assert closure_rc(good) != 0 and b"null" in vs.lib.vs_last_error()                      # valid shape, null buffers
    assert vs.lib.vs_u8_to_f32(None, None, 16, None) != 0


def test_bench_clock_sampler_filters_the_timed_region(monkeypatch):
    """bench.py's clocks object: only nvidia-smi samples taken inside the timed region count; a region shorter than one sampling
    period falls back to the samples of the warm-up just before it; throttle reasons are collected from the kept rows."""
    import time
    import bench

    class FakeProc:
        stdout = ()
        def terminate(self): pass
        def wait(self, timeout=None): return 0

    def row(sm, power_cap="Not Active", hw="Not Active"):
        return ["0", str(sm), "1965", "700.0", "0x4", hw, "Not Active", "Not Active", power_cap]

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    now = time.perf_counter()
    s.rows = [(now - 3.0, row(300)), (now - 0.2, row(1700, power_cap="Active")), (now + 0.01, row(1800)), (now + 0.02, row(1900, power_cap="Active"))]
    s.t0 = now
    time.sleep(0.05)
    out = s.stop()
    assert out["samples_in_timed_region"] == 2 and out["sm_mhz"] == 1850.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    s2.proc = FakeProc()
    now = time.perf_counter()
    s2.rows = [(now - 3.0, row(300, hw="Active")), (now - 0.2, row(1700))]
    s2.t0 = now
    out = s2.stop()
    assert out["samples_in_timed_region"] == 0 and out["samples"] == 1 and out["sm_mhz"] == 1700.0 and out["reasons"] == []
    s3 = bench.ClockSampler(0)                  # nvidia-smi missing: reported, not fatal
    assert s3.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_bench_reference_arm_prints_one_json_line_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): runs without a GPU and without the CUDA library,
    prints ONE JSON line with the contract's keys, and reports the neuron sample it actually ran with its extrapolation."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--trials", "16",
                        "--trials-test", "8", "--features", "256", "--neurons", "16", "--ref-budget", "2"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training video frames/sec" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["dtype"] == "f64"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb, cfg = d["cpu_baseline"], d["config"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert 1 <= cfg["neurons"] <= cfg["neurons_full"] == 16
    ext = cfg["extrapolation"]
    assert ext["sampled"] == "neurons" and ext["run"] == cfg["neurons"] and 0 < ext["factor_on_value"] <= 1.0
    assert cfg["extrapolated"] == (cfg["neurons"] < 16)


def test_rrr_init_stream_marks_reproduce_the_sequential_draw(monkeypatch):
    """RRRGD(init_plan=...) on a rank that owns ONE session of a joint model: jumping over the foreign sessions with the
    remembered stream positions gives bit-identical U, V and numpy global state to drawing and dropping them
    (src/model/rrr.py:35-49: one global stream, U then V per session, the last V kept)."""
    import model.rrr as mr
    rng = np.random.default_rng(0)
    plan = [("a", 5, 40, 7), ("b", 3, 33, 7), ("c", 4, 21, 7), ("d", 6, 18, 7)]

    def td(eid):
        _, N, C, T = next(p for p in plan if p[0] == eid)
        return {eid: {"X": [rng.standard_normal((6, T, C)), rng.standard_normal((3, T, C))],
                      "y": [rng.standard_normal((6, T, N)), rng.standard_normal((3, T, N))]}}

    for own in ("a", "b", "c", "d"):
        data = td(own)
        monkeypatch.setenv("VS_RRR_STREAM_MARKS", "0")
        ref = mr.RRRGD(data, 3, l2=1.0, init_plan=plan)
        st_ref = np.random.get_state()
        monkeypatch.setenv("VS_RRR_STREAM_MARKS", "1")
        for _ in range(2):                       # first pass may record the marks, the second one uses them
            got = mr.RRRGD(data, 3, l2=1.0, init_plan=plan)
            st = np.random.get_state()
            for k in ref.model:
                assert torch.equal(ref.model[k], got.model[k]), (own, k)
            assert st[2:] == st_ref[2:] and np.array_equal(st[1], st_ref[1])
    assert any(k[2] == "U" for k in mr._STREAM_MARKS) and any(k[2] == "V" for k in mr._STREAM_MARKS)
    # the whole joint model in one process equals the reference's own stream for the first session
    np.random.seed(0)
    U0 = np.random.normal(size=(5, 39, 3)) / np.sqrt(7 * 3)
    full = mr.RRRGD({**td("a"), **td("b"), **td("c"), **td("d")}, 3, l2=1.0)
    assert np.array_equal(full.model["a_U"].detach().numpy(), U0)


def test_neuron_group_bounds_cover_wide_sessions():
    """model.rrr._neuron_group_bounds: contiguous ranges of at most 160 neurons covering 0..N, one range up to 160."""
    from model.rrr import _neuron_group_bounds
    for N in (1, 16, 144, 160, 161, 300, 436, 450, 1024):
        b = _neuron_group_bounds(N)
        assert b[0][0] == 0 and b[-1][1] == N
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(0 < n1 - n0 <= 160 for n0, n1 in b)
        assert (len(b) == 1) == (N <= 160)
        assert all(n0 % 16 == 0 for n0, _ in b)                # group starts keep U / dU slices 16-byte aligned
    assert _neuron_group_bounds(436) == [(0, 112), (112, 224), (224, 336), (336, 436)]
