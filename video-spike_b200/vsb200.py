"""ctypes binding of libvs_b200.so (include/vs_b200.h) -- the only door from the Python host
code into the sm_100a kernels.  PyTorch supplies device memory and streams; every pointer handed
to the library is `tensor.data_ptr()` of a CUDA tensor owned by the caller.

There is NO CPU fallback: importing this module without the built library raises, and every
wrapper raises if a tensor is not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvs_b200.so")

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
OPERAND_BF16, OPERAND_F16 = 0, 1
RRR_MODE_CLASSIC, RRR_MODE_EXACT, RRR_MODE_DENSE = 0, 1, 2
MAX_LAYERS = 16


class VsError(RuntimeError):
    pass


class AdamWHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("step", C.c_int64)]


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int64 * (MAX_LAYERS + 1)), ("relu", C.c_int32 * MAX_LAYERS)] + \
               [(name, C.c_void_p * MAX_LAYERS) for name in
                ("W", "b", "mW", "vW", "mb", "vb", "act", "gact", "gW", "gb")]


LBFGS_MAX_HIST = 100


class LbfgsDev(C.Structure):
    """vs_lbfgs_dev: the device-resident state of the device-driven L-BFGS (include/vs_b200.h)."""
    _fields_ = [("m", C.c_int32), ("n_iter", C.c_int32), ("total_iter", C.c_int32), ("func_evals", C.c_int32),
                ("cur_evals", C.c_int32), ("done", C.c_int32), ("have_prev", C.c_int32), ("have_s", C.c_int32),
                ("s_cur", C.c_int32), ("y_next", C.c_int32), ("n_free", C.c_int32), ("pad0", C.c_int32),
                ("free_slots", C.c_int32 * (2 * LBFGS_MAX_HIST + 8)),
                ("s_slots", C.c_int32 * LBFGS_MAX_HIST), ("y_slots", C.c_int32 * LBFGS_MAX_HIST),
                ("H_diag", C.c_double), ("prev_loss", C.c_double), ("loss", C.c_double), ("t", C.c_double),
                ("gtd", C.c_double), ("dmax", C.c_double),
                ("coef", C.c_double * (2 * LBFGS_MAX_HIST + 2)),
                ("out", C.c_double * (8 + 6 * LBFGS_MAX_HIST)),
                ("SY", C.c_double * (LBFGS_MAX_HIST * LBFGS_MAX_HIST)), ("YY", C.c_double * (LBFGS_MAX_HIST * LBFGS_MAX_HIST))]


LBFGS_CMAX = 64


class LbfgsCDev(C.Structure):
    """vs_lbfgs_cdev: device-resident state of the compact-history L-BFGS (include/vs_b200.h)."""
    _fields_ = [("m", C.c_int32), ("nb", C.c_int32), ("n_iter", C.c_int32), ("total_iter", C.c_int32), ("func_evals", C.c_int32),
                ("cur_evals", C.c_int32), ("done", C.c_int32), ("have_prev", C.c_int32), ("have_s", C.c_int32), ("pad0", C.c_int32),
                ("iy", C.c_int32 * LBFGS_CMAX),
                ("H_diag", C.c_double), ("prev_loss", C.c_double), ("loss", C.c_double), ("t", C.c_double), ("gtd", C.c_double),
                ("dmax", C.c_double),
                ("coef", C.c_double * (LBFGS_CMAX + 2)), ("Acur", C.c_double * LBFGS_CMAX), ("out", C.c_double * (8 + 3 * LBFGS_CMAX)),
                ("P", C.c_double * (LBFGS_CMAX * LBFGS_CMAX)), ("A", C.c_double * (LBFGS_CMAX * LBFGS_CMAX)),
                ("SY", C.c_double * (LBFGS_CMAX * LBFGS_CMAX)), ("YY", C.c_double * (LBFGS_CMAX * LBFGS_CMAX))]


class RrrDims(C.Structure):
    _fields_ = [("K", C.c_int64), ("T", C.c_int64), ("C1", C.c_int64), ("N", C.c_int64), ("r", C.c_int64),
                ("planes", C.c_int32), ("ldc", C.c_int64), ("ldr", C.c_int64), ("fmt", C.c_int32), ("mode", C.c_int32)]


class RrrExactOps(C.Structure):
    """vs_rrr_exact_ops: device pointers of the exact-operand tables of one split (include/vs_b200.h)."""
    _fields_ = [("Xi", C.c_void_p), ("isdT", C.c_void_p), ("qT", C.c_void_p), ("ldt", C.c_int64), ("y_lo", C.c_void_p),
                ("Xc", C.c_void_p), ("isd", C.c_void_p), ("qh", C.c_void_p), ("isdmax", C.c_void_p)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C video-spike_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, sz, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_double
    sig = {
        "vs_version": (C.c_int, []),
        "vs_last_error": (C.c_char_p, []),
        "vs_device_ok": (C.c_int, []),
        "vs_u8_to_f32": (C.c_int, [vp, vp, i64, vp]),
        "vs_h2d_select_frames": (C.c_int, [vp, i64, i64, i64, vp, i64, vp, vp]),
        "vs_u8_to_bf16": (C.c_int, [vp, vp, i64, vp]),
        "vs_gather_windows": (C.c_int, [vp, i64, i64, vp, i64, i64, vp, vp]),
        "vs_linear_fwd_workspace": (sz, [i64, i64, i64]),
        "vs_linear_fwd": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, vp, sz, vp]),
        "vs_linear_bwd_workspace": (sz, [i64, i64, i64]),
        "vs_linear_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, vp, sz, vp]),
        "vs_poisson_nll": (C.c_int, [vp, vp, vp, vp, i64, vp]),
        "vs_bits_per_spike": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
        "vs_r2_rows": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
        "vs_adamw": (C.c_int, [vp, vp, vp, vp, i64, AdamWHyper, vp]),
        "vs_dw_adamw_fused": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, AdamWHyper, vp]),
        "vs_mlp_workspace": (sz, [C.POINTER(Mlp), i64]),
        "vs_mlp_train_step": (C.c_int, [C.POINTER(Mlp), vp, vp, vp, i64, AdamWHyper, vp, i32, vp, sz, vp]),
        "vs_mlp_train_step_rowpar": (C.c_int, [C.POINTER(Mlp), vp, vp, vp, i64, AdamWHyper, vp, i32, vp, sz, vp, i32]),
        "vs_mlp_forward": (C.c_int, [C.POINTER(Mlp), vp, vp, vp, i64, vp, i32, vp, sz, vp]),
        "vs_rrr_ldc": (i64, [i64]),
        "vs_rrr_ldr": (i64, [i64, i64]),
        "vs_rrr_pack": (C.c_int, [vp, i64, i64, RrrDims, vp, vp, vp, vp, vp]),
        "vs_rrr_colstats": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "vs_rrr_pack_u8": (C.c_int, [vp, i64, vp, vp, vp, RrrDims, vp, vp, vp, vp, vp]),
        "vs_rrr_smooth_y": (C.c_int, [vp, i64, i64, i64, dbl, vp, vp, vp, vp]),
        "vs_rrr_smooth_y2": (C.c_int, [vp, i64, i64, i64, dbl, vp, vp, vp, vp, vp]),
        "vs_rrr_ldt": (i64, [i64]),
        "vs_rrr_exact_supported": (C.c_int, [i64, i64, i64, i64, i64]),
        "vs_rrr_pack_u8_exact": (C.c_int, [vp, i64, vp, vp, vp, RrrDims, vp, C.POINTER(RrrExactOps), vp, vp, vp]),
        "vs_rrr_pack_u8_fused": (C.c_int, [vp, i64, vp, vp, vp, i32, RrrDims, vp, C.POINTER(RrrExactOps), vp, vp, vp]),
        "vs_rrr_closure_exact": (C.c_int, [RrrDims, vp, C.POINTER(RrrExactOps), vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, sz, vp]),
        "vs_rrr_predict_exact": (C.c_int, [RrrDims, C.POINTER(RrrExactOps), vp, vp, vp, vp, vp, vp, sz, vp]),
        "vs_colstats_f32": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "vs_rrr_workspace": (sz, [RrrDims]),
        "vs_rrr_closure": (C.c_int, [RrrDims, vp, vp, vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, i32, vp, sz, vp]),
        "vs_rrr_predict": (C.c_int, [RrrDims, vp, vp, vp, vp, vp, vp, i32, vp, sz, vp]),
        "vs_lbfgs_workspace": (sz, [i64, i32]),
        "vs_lbfgs_dots": (C.c_int, [i64, vp, vp, vp, vp, vp, i64, i32, vp, vp, i32, vp, vp, sz, vp]),
        "vs_lbfgs_direction": (C.c_int, [i64, vp, vp, i64, i32, vp, vp, i32, vp, dbl, vp, vp, vp, vp]),
        "vs_lbfgs_dev_init_host": (C.c_int, [C.POINTER(LbfgsDev), i32]),
        "vs_lbfgs_dev_state_bytes": (sz, []),
        "vs_lbfgs_dev_workspace": (sz, [i64]),
        "vs_lbfgs_dev_dots": (C.c_int, [vp, i64, vp, vp, vp, i64, i32, vp, sz, vp]),
        "vs_lbfgs_dev_update": (C.c_int, [vp, vp, dbl, dbl, dbl, i32, i32, i32, vp]),
        "vs_lbfgs_dev_direction": (C.c_int, [vp, i64, vp, vp, i64, i32, vp, vp]),
        "vs_lbfgs_cdev_init_host": (C.c_int, [C.POINTER(LbfgsCDev)]),
        "vs_lbfgs_cdev_state_bytes": (sz, []),
        "vs_lbfgs_cdev_workspace": (sz, [i64]),
        "vs_lbfgs_cdev_dots": (C.c_int, [vp, i64, vp, vp, vp, i64, vp, sz, vp]),
        "vs_lbfgs_cdev_update": (C.c_int, [vp, vp, dbl, dbl, dbl, i32, i32, vp]),
        "vs_lbfgs_cdev_direction": (C.c_int, [vp, i64, vp, vp, i64, vp, vp]),
        "vs_host_lbfgs_two_loop": (C.c_int, [i32, dbl, vp, vp, vp, vp, i32, dbl, vp, vp]),
        "vs_host_rng_seed": (C.c_int, [vp, C.c_uint32]),
        "vs_host_rng_normal": (C.c_int, [vp, i64, dbl, vp, i32]),
        "vs_host_rng_get_state": (C.c_int, [vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
        "vs_host_rng_set_state": (C.c_int, [vp, vp, C.c_int32, C.c_int32, dbl]),
        "vs_gemm_tn": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, i64, i64, i32, i32, vp]),
        "vs_launch_count": (i64, []),
        "vs_launch_count_reset": (None, []),
        "vs_profile_enable": (None, [i32]),
        "vs_profile_read": (i64, [i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sig.keys())


lib, EXPORTS = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise VsError(f"libvs_b200 error {rc}: {lib.vs_last_error().decode()}")


def ptr(t):
    """Raw device pointer of a CUDA tensor (None -> NULL).  Refuses CPU tensors: no fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VsError("libvs_b200 only takes CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise VsError("libvs_b200 needs contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_b200() -> None:
    if not torch.cuda.is_available():
        raise VsError("no CUDA device: video-spike_b200 has no CPU path")
    if not lib.vs_device_ok():
        raise VsError("video-spike_b200 kernels are built for sm_100a (B200) only")


# ---------------------------------------------------------------- thin typed wrappers
def u8_to_f32(frames: torch.Tensor) -> torch.Tensor:
    """src/loader/base.py:39,54 -- uint8 frames -> float32, values stay 0..255."""
    assert frames.dtype == torch.uint8
    out = torch.empty(frames.shape, dtype=torch.float32, device=frames.device)
    if frames.numel():
        check(lib.vs_u8_to_f32(ptr(frames), ptr(out), frames.numel(), stream()))
    return out


def u8_to_bf16(frames: torch.Tensor) -> torch.Tensor:
    assert frames.dtype == torch.uint8
    out = torch.empty(frames.shape, dtype=torch.bfloat16, device=frames.device)
    if frames.numel():
        check(lib.vs_u8_to_bf16(ptr(frames), ptr(out), frames.numel(), stream()))
    return out


def poisson_nll(logits: torch.Tensor, target: torch.Tensor, want_grad: bool = True):
    """Returns (loss_sum double[1], dlogits or None); mean loss = loss_sum / numel."""
    loss = torch.empty(1, dtype=torch.float64, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    check(lib.vs_poisson_nll(ptr(logits), ptr(target), ptr(loss), ptr(dl), logits.numel(), stream()))
    return loss, dl


def adamw(p, g, m, v, hyper: AdamWHyper) -> None:
    check(lib.vs_adamw(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), hyper, stream()))


def gemm_tn(A: torch.Tensor, B: torch.Tensor, engine: int = ENGINE_AUTO) -> torch.Tensor:
    """C = A @ B.T for K-major A (M,K), B (N,K); bf16 or fp32(tf32) operands.  Test hook."""
    assert A.dtype == B.dtype and A.dtype in (torch.bfloat16, torch.float32, torch.float16)
    M, K = A.shape
    N = B.shape[0]
    Cm = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(lib.vs_gemm_tn(ptr(A), ptr(B), ptr(Cm), M, N, K, A.stride(0), B.stride(0), N,
                         {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}[A.dtype], engine, stream()))
    return Cm


def profile_read(tag: int):
    """(count, total_ms, min_ms, max_ms) of the event-bracketed launches of one kernel class."""
    tot, mn, mx = C.c_double(), C.c_double(), C.c_double()
    cnt = lib.vs_profile_read(tag, C.byref(tot), C.byref(mn), C.byref(mx))
    return int(cnt), tot.value, mn.value, mx.value


class LegacyNormalStream:
    """numpy's legacy global stream (`np.random.seed(s)`; `np.random.normal(size=...)`) reproduced bit for bit
    by the multi-threaded host generator of libvs_b200 (csrc/host_rng.cpp).  Used by RRRGD.__init__
    (src/model/rrr.py:35,42-43 draws ~8 M normals per session; numpy needs ~0.2 s for that)."""
    STATE_BYTES = 2560

    def __init__(self, seed: int):
        self._state = C.create_string_buffer(self.STATE_BYTES)
        check(lib.vs_host_rng_seed(self._state, int(seed) & 0xFFFFFFFF))

    def normal(self, size, divisor: float = 1.0, threads: int = 0, out=None):
        """`out`: optional C-contiguous float64 array / CPU tensor of the right size to fill (e.g. pinned memory)."""
        import numpy as np
        if out is None:
            out = np.empty(size, dtype=np.float64)
            ptr, count = out.ctypes.data, out.size
        else:
            ptr, count = (out.data_ptr(), out.numel()) if isinstance(out, torch.Tensor) else (out.ctypes.data, out.size)
            if count != int(np.prod(size)):
                raise VsError("LegacyNormalStream.normal: `out` has the wrong number of elements")
        check(lib.vs_host_rng_normal(self._state, count, float(divisor), C.c_void_p(ptr), int(threads)))
        return out

    def snapshot(self) -> bytes:
        """The generator's whole state (position in the stream), e.g. to resume later from this point."""
        return bytes(self._state.raw)

    def restore(self, blob: bytes):
        if len(blob) != self.STATE_BYTES:
            raise VsError("LegacyNormalStream.restore: wrong state size")
        C.memmove(self._state, blob, self.STATE_BYTES)

    def export_to_numpy(self):
        """Leave numpy's GLOBAL RandomState exactly where the reference's draws would have left it."""
        import numpy as np
        key = np.empty(624, dtype=np.uint32)
        pos, has, cached = C.c_int32(), C.c_int32(), C.c_double()
        check(lib.vs_host_rng_get_state(self._state, key.ctypes.data_as(C.c_void_p), C.byref(pos), C.byref(has), C.byref(cached)))
        np.random.set_state(("MT19937", key, int(pos.value), int(has.value), float(cached.value)))
