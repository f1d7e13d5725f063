"""Reduced-rank regression: drop-in for the reference's src/model/rrr.py.

Same public surface -- `RRRGD(train_data, ncomp, l2)`, `.model` (an nn.ParameterDict with keys
"{eid}_U" (N,C-1,r), "{eid}_b" (N,1,T), "V" (r,T), float64, initialised under np.random.seed(0)
exactly as rrr.py:35-49), `predict_y`, `predict_y_fr`, `compute_MSE_RRRGD`, `regression_loss`,
`state_dict` (rrr.py:64-71), `train_model`, `train_model_main` -- and the same optimiser: ONE
`torch.optim.LBFGS(...).step(closure)` (rrr.py:177,199).

What changed is the closure.  The reference rebuilds beta = cat(U@V, b) twice, re-uploads X and y
on every evaluation and differentiates an einsum with autograd (rrr.py:122-155).  Here each split
of each session is packed ONCE into device-resident bf16 operand planes (vs_rrr_pack) and every
closure evaluation is one vs_rrr_closure call: two tcgen05 GEMMs plus ordered reductions that
return the loss and write dU, dV, db straight into `param.grad`.  No CPU path exists.

Precision (DESIGN.md "RRR precision"): the reference's un-line-searched L-BFGS amplifies operand rounding by 3-4
orders of magnitude, so the defaults are the modes whose WHOLE FIT stays within 1e-3 of the float64 reference:
splits packed from uint8 frames use the exact-operand mode (`rrr_mode`), float64 numpy splits 3 bf16 residual
planes (`planes`, env VS_RRR_PLANES).  planes = 1 is the fastest, non-parity setting.  Accumulation is fp32 in
TMEM, reductions fp64.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn
from torch import optim

import vsb200 as vs
from optim import FusedLBFGS


def operand_format(planes, operand=None):
    """16-bit format of the tensor-core operand planes.  "bf16" (default; fp32 range, 8 significant bits) or "f16"
    (IEEE half: 11 significant bits at the same tcgen05 rate, so every closure evaluation is ~8x closer to float64, but
    |x| <= 65504 -- the exploding test features of SURVEY A18 overflow it, which the pack calls report loudly).
    Select with `operand=` or VS_RRR_OPERAND."""
    operand = operand or os.environ.get("VS_RRR_OPERAND") or "bf16"
    if operand not in ("bf16", "f16"):
        raise ValueError("operand must be 'bf16' or 'f16'")
    return vs.OPERAND_F16 if operand == "f16" else vs.OPERAND_BF16


def _op_dtype(fmt):
    return torch.float16 if fmt == vs.OPERAND_F16 else torch.bfloat16


def rrr_mode(mode=None):
    """"exact" (default) or "classic" -- how splits packed from uint8 frames represent X (include/vs_b200.h,
    VS_RRR_MODE_*).  exact: the mode whose whole fit lands within 1e-3 of the float64 reference (z-score as hi + lo half
    planes for the forward, exact integer frames for the backward, float64 epilogues, float64 L-BFGS history).
    classic: `planes` residual planes of the z-scored matrix for both contractions (planes = 1 is the fastest setting;
    its fit is ~2e-3 away from the reference's, DESIGN.md "RRR precision").  Select with `mode=` or VS_RRR_MODE."""
    mode = mode or os.environ.get("VS_RRR_MODE") or "exact"
    if mode not in ("exact", "dense", "classic"):
        raise ValueError("mode must be 'exact', 'dense' or 'classic'")
    return mode


def np2tensor(v):
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def np2param(v, grad=True):
    return nn.Parameter(np2tensor(v), requires_grad=grad)


def tensor2np(v):
    return v.numpy()


def get_device():
    """rrr.py:16-26, minus the CPU branch: this implementation only runs on a B200."""
    vs.require_b200()
    print("GPU is available")
    return torch.device("cuda")


_SIDE_STREAMS = {}
_PINNED_STAGES = {}


def _neuron_group_bounds(N, limit=160):
    """Neuron ranges of at most `limit` (sizes a multiple of 16 where possible) covering 0..N."""
    if N <= limit:
        return [(0, N)]
    ng = -(-N // 144)
    size = -(-(-(-N // ng)) // 16) * 16
    return [(g, min(N, g + size)) for g in range(0, N, size)]


# positions of the init stream (LegacyNormalStream.snapshot) after a given sequence of session draws from seed 0:
# key (shapes of the sessions drawn so far, n_comp, "U" | "V") -- see RRRGD.__init__
_STREAM_MARKS = {}


def _stream_cache_enabled():
    if len(_STREAM_MARKS) > 4096:
        _STREAM_MARKS.clear()
    return os.environ.get("VS_RRR_STREAM_MARKS", "1") != "0"


def _pinned_stage(numel):
    """A reusable pinned float64 staging buffer of `numel` elements (allocating pinned memory costs milliseconds).  ONE
    grow-only buffer per process, sliced to the size asked for: a sweep over sessions of different sizes pins the largest
    one, not one buffer per size.  The event guards reuse: the previous upload out of the buffer must have finished."""
    st = _PINNED_STAGES.get("stage")
    if st is not None:
        st["event"].synchronize()
    if st is None or st["full"].numel() < numel:
        st = _PINNED_STAGES["stage"] = {"full": torch.empty(numel, dtype=torch.float64).pin_memory(), "event": torch.cuda.Event()}
    st["buf"] = st["full"][:numel]
    return st



def _side_stream(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class _LazyY:
    """The z-scored targets of a split that is still being produced on the side stream: behaves like the tensor once
    touched (joins the stream first).  Only `shape` is available without joining (RRRGD.__init__ reads it)."""

    def __init__(self, split):
        self._split = split

    @property
    def shape(self):
        return self._split.y.shape

    def tensor(self):
        self._split.wait_ready()
        return self._split.y

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.tensor(), name)

    def __getitem__(self, i):
        return self.tensor()[i]


class _PackedSplit:
    """Device-resident operands of one (session, split): built once, reused by every closure."""

    CHUNK_BYTES = 1 << 30   # fp64 staging buffer bound for the host->device upload

    def __init__(self, X, y, r, planes, device, fmt=0):
        X = np.ascontiguousarray(X)
        K, T, C = X.shape
        N = y.shape[2]
        self.K, self.T, self.C1, self.N = K, T, C - 1, N
        d = vs.RrrDims(K, T, C - 1, N, r, planes, vs.lib.vs_rrr_ldc(C - 1), vs.lib.vs_rrr_ldr(K, T), fmt)
        self.dims = d
        KT = K * T
        self.exact = None
        self.Xa = torch.empty((planes, KT, d.ldc), dtype=_op_dtype(fmt), device=device)
        self.Xb = torch.empty((planes, C - 1, d.ldr), dtype=_op_dtype(fmt), device=device)
        self.xl = torch.empty(KT, dtype=torch.float32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        trials_per_chunk = max(1, min(K, self.CHUNK_BYTES // (8 * C * T)))
        st = vs.stream()
        for k0 in range(0, K, trials_per_chunk):
            k1 = min(K, k0 + trials_per_chunk)
            chunk = torch.from_numpy(X[k0:k1]).to(device=device, dtype=torch.float64).contiguous()
            vs.check(vs.lib.vs_rrr_pack(vs.ptr(chunk), k0, k1 - k0, d, vs.ptr(self.Xa), vs.ptr(self.Xb), vs.ptr(self.xl),
                                        vs.ptr(self.overflow), st))
            del chunk
        self.y = torch.from_numpy(np.ascontiguousarray(y)).to(device=device, dtype=torch.float32).contiguous()

    @classmethod
    def from_device(cls, dims, Xa, Xb, xl, y, overflow=None, exact=None):
        """Wrap operands that were produced on the device (vs_rrr_pack_u8 / vs_rrr_pack_u8_exact path).
        `exact`: {"isdT", "qT", "ldt", "y_lo"} of an exact-operand split (Xb then holds the integer operand, or None)."""
        self = cls.__new__(cls)
        self.K, self.T, self.C1, self.N = dims.K, dims.T, dims.C1, dims.N
        self.dims, self.Xa, self.Xb, self.xl, self.y, self.overflow = dims, Xa, Xb, xl, y, overflow
        self.exact = exact
        self.ready = None
        self._keep = None
        return self

    def exact_ops(self, y_lo=None):
        """ctypes vs_rrr_exact_ops of this split (the tensors stay referenced by self.exact); `y_lo`: a neuron group's own."""
        ex = self.exact
        g = lambda k: vs.ptr(ex.get(k)) if ex.get(k) is not None else None      # noqa: E731
        return vs.RrrExactOps(vs.ptr(self.Xb) if self.Xb is not None else None, g("isdT"), g("qT"), int(ex["ldt"]),
                              vs.ptr(y_lo) if y_lo is not None else g("y_lo"), g("Xc"), g("isd"), g("qh"), g("isdmax"))

    def neuron_groups(self):
        """[(n0, n1, dims, y, y_lo)] -- the exact-operand kernels hold at most 160 neurons per launch; wider sessions
        (BASELINE: N = 436) are evaluated group by group: the model is separable over neurons given V (loss, the l2 term and
        dV are sums over neurons; U, b, dU, db of a group are contiguous slices)."""
        ex = self.exact or {}
        return ex.get("groups") or [(0, self.N, self.dims, self.y, ex.get("y_lo"))]

    def __del__(self):
        ev = getattr(self, "ready", None)
        if ev is not None:
            try:
                ev.synchronize()
            except Exception:
                pass
        self._keep = None

    def wait_ready(self):
        """Join the side stream that produced this split (pack_session_from_frames) before the first read."""
        ev = getattr(self, "ready", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self.ready = None
            # main-stream tensors the side-stream kernels read (frame indices, z-score statistics): the caching allocator
            # may hand their blocks to the next main-stream allocation the moment they die, so they live exactly until
            # the main stream is ordered behind the side stream
            self._keep = None

    def check_range(self):
        """Half-precision operands: fail loudly if a value left the half range while packing (checked once, lazily,
        so the pack stays asynchronous)."""
        if self.overflow is not None:
            flag, self.overflow = int(self.overflow.item()), None
            if flag:
                raise vs.VsError("RRR operands exceed the IEEE-half range (e.g. features that are constant in the train split, "
                                 "SURVEY A18): use operand='bf16' / VS_RRR_OPERAND=bf16")

    @property
    def shape(self):
        """(K, T, ncoef) of the matrix this split stands for (ncoef counts the bias column)."""
        return (self.K, self.T, self.C1 + 1)


class RRRGD():
    def __init__(self, train_data, ncomp, l2=0., planes=None, engine=None, init_plan=None, operand=None, device=None, draw=True):
        """`init_plan` (session-sharded joint model, parallel.py): [(eid, N, ncoef, T)] of ALL sessions of the joint
        model in the reference's iteration order.  The init stream is drawn for every session in that order so that
        this rank's U_s and the shared V are bit-identical to the single-process joint model; sessions that are not
        in `train_data` are drawn and dropped.
        `device`: create the parameters directly on that CUDA device -- the init stream is generated into a pinned
        staging buffer and uploaded asynchronously, instead of CPU parameters + a later blocking `.to(device)`.
        `draw=False` (replicated models, parallel.build_trial_sharded_model): U and V are left unset -- another rank draws
        them and broadcasts; numpy's global state is still moved to where the draws would have left it when that position
        is known (_STREAM_MARKS)."""
        self.l2 = l2
        self.eids = list(train_data.keys())
        self.withbias = True
        # host (float64 numpy) splits are packed into this many residual planes; the default 3 keeps the whole fit within
        # 1e-3 of the float64 reference (1 = fastest, ~2e-3 away; device-packed splits carry their own layout)
        self.planes = int(planes if planes is not None else os.environ.get("VS_RRR_PLANES", "3"))
        self.engine = int(engine if engine is not None else os.environ.get("VS_ENGINE", str(vs.ENGINE_AUTO)))
        self.fmt = operand_format(self.planes, operand)
        # splits packed in the exact-operand mode (pack_session_from_frames) carry their own layout: two half planes
        self.exact = any(isinstance(x, _PackedSplit) and x.dims.mode != vs.RRR_MODE_CLASSIC
                         for e in train_data.values() for x in e["X"])
        if self.exact:
            self.planes, self.fmt = 2, vs.OPERAND_F16

        # rrr.py:35 seeds numpy's GLOBAL legacy stream with 0 regardless of the user's seed and draws
        # U then V per session from it.  The same stream -- bit for bit -- comes from the multi-threaded host
        # generator of libvs_b200 (numpy's scalar loop costs more than the whole GPU fit); numpy's global state
        # is then left exactly where the reference would leave it.
        rng = vs.LegacyNormalStream(0)
        self.N = 0
        params = {}
        V = None
        if init_plan is None:
            init_plan = [(eid, train_data[eid]['y'][0].shape[2], train_data[eid]['X'][0].shape[2], train_data[eid]['y'][0].shape[1])
                         for eid in train_data]
        self.eids = [eid for eid, *_ in init_plan if eid in train_data]
        # Sessions of the plan that live on OTHER ranks only advance the stream.  The stream position after a given sequence
        # of draws is a pure function of seed 0 and the shapes, so it is remembered (2.5 KB per position, _STREAM_MARKS) the
        # first time it is reached: later models with the same plan jump over foreign sessions instead of drawing and
        # dropping ~8 M normals for each of them (on n ranks that was n x the host work, all ranks competing for the same
        # cores).  A rank's own U and the V that is kept are always drawn.
        shapes = tuple((int(N), int(ncoef), int(T)) for _, N, ncoef, T in init_plan)
        last = len(init_plan) - 1
        if not draw:
            for eid, N, ncoef, T in init_plan:
                if eid not in train_data:
                    continue
                _y = train_data[eid]['y'][0]
                b = (_y.double().mean(0).T.unsqueeze(1).contiguous() if isinstance(_y, torch.Tensor)
                     else np.ascontiguousarray(np.expand_dims(_y.mean(0).T, 1)))
                U = (torch.empty((N, ncoef - 1, ncomp), dtype=torch.float64, device=device) if device is not None
                     else np.empty((N, ncoef - 1, ncomp)))
                params[f"{eid}_U"] = np2param(U)
                params[f"{eid}_b"] = np2param(b)
                self.N += N
            V = np.zeros((ncomp, init_plan[last][3]))
            mark = _STREAM_MARKS.get((shapes, int(ncomp), "V"))
            if mark is not None:
                rng.restore(mark)
            else:
                rng = None                                       # position unknown: numpy's global state is left alone
            init_plan = []                                       # nothing is drawn below
        for i, (eid, N, ncoef, T) in enumerate(init_plan):
            scale = float(np.sqrt(T * ncomp))
            own = eid in train_data
            if not own:
                # the furthest remembered position up to the next draw that is needed: the next own session's U, or the last V
                nxt = next((j for j in range(i + 1, len(init_plan)) if init_plan[j][0] in train_data), None)
                target = (nxt - 1, "V") if nxt is not None else (last, "U")
                mark = _STREAM_MARKS.get((shapes[:target[0] + 1], int(ncomp), target[1])) if target[0] >= i else None
                if mark is not None and _stream_cache_enabled():
                    rng.restore(mark)
                    if nxt is None:                             # positioned after the last session's U: its V is the one kept
                        V = rng.normal((ncomp, init_plan[last][3]), float(np.sqrt(init_plan[last][3] * ncomp)))
                        _STREAM_MARKS[(shapes, int(ncomp), "V")] = rng.snapshot()
                        break
                    continue                                     # positioned after session nxt - 1: the loop resumes at nxt
            if device is not None and own:
                stage = _pinned_stage(N * (ncoef - 1) * ncomp)
                rng.normal((N, ncoef - 1, ncomp), scale, out=stage["buf"])
                U = torch.empty((N, ncoef - 1, ncomp), dtype=torch.float64, device=device)
                U.copy_(stage["buf"].view(N, ncoef - 1, ncomp), non_blocking=True)
                stage["event"].record(torch.cuda.current_stream(device))
            else:
                U = rng.normal((N, ncoef - 1, ncomp), scale)
            _STREAM_MARKS.setdefault((shapes[:i + 1], int(ncomp), "U"), rng.snapshot())
            V = rng.normal((ncomp, T), scale)                  # redrawn per eid, the last one is kept (rrr.py:43,49)
            _STREAM_MARKS.setdefault((shapes[:i + 1], int(ncomp), "V"), rng.snapshot())
            if not own:
                continue
            _y = train_data[eid]['y'][0]       # (K, T, N)
            if isinstance(_y, torch.Tensor):   # targets already on the device (pack_session_from_frames)
                b = _y.double().mean(0).T.unsqueeze(1).contiguous()
            else:
                b = np.ascontiguousarray(np.expand_dims(_y.mean(0).T, 1))
            params[f"{eid}_U"] = np2param(U)
            params[f"{eid}_b"] = np2param(b)
            self.N += N
        if rng is not None:
            rng.export_to_numpy()
        params['V'] = np2param(V)
        self.n_comp, self.T = params['V'].shape
        self.model = nn.ParameterDict(params)
        if device is not None:
            self.model.to(device)                 # U is already there; V and b are a few KB
        self._packed = {}
        self._ws = None
        self.n_closure_evals = 0

    def make_optimizer(self, history_dtype=None, device_driven=None):
        """The optimiser `train_model_main` builds (rrr.py:199: `optim.LBFGS(params)` with torch's defaults).
        Curvature-pair storage follows the operand precision: float32 with plain-bf16 operands (planes == 1, where a
        closure evaluation carries ~1e-3 of rounding anyway), float64 with residual planes (parity mode).  Override
        with `history_dtype` or the environment variable VS_LBFGS_HIST=f32|f64."""
        if history_dtype is None:
            env = os.environ.get("VS_LBFGS_HIST")
            if env:
                history_dtype = torch.float32 if env.lower() in ("f32", "float32", "fp32") else torch.float64
            else:
                history_dtype = torch.float32 if self.planes == 1 else torch.float64
        if device_driven is None:     # decisions on the device, no host sync per iteration (optim.FusedLBFGS); VS_LBFGS_DEVICE=0 disables
            device_driven = os.environ.get("VS_LBFGS_DEVICE", "1") != "0"
        # float64 history: one vector per evaluation instead of a pair (half the traffic of the vector passes); VS_LBFGS_COMPACT=0 disables
        compact = bool(device_driven and history_dtype == torch.float64 and os.environ.get("VS_LBFGS_COMPACT", "1") != "0")
        opt = FusedLBFGS(self.model.parameters(), history_dtype=history_dtype, device_driven=device_driven, compact=compact)
        opt.closure_overwrites_grads = True     # loss_and_grad writes every gradient element (see there): zero_grad() need not memset
        return opt

    def train(self):
        self.model.train()

    def eval(self):
        self.model.eval()

    def to(self, device):
        self.model.to(device)

    def state_dict(self):
        return {"model": {k: v.cpu() for k, v in self.model.state_dict().items()},
                "l2": self.l2, "eids": self.eids, "N": self.N, "T": self.T, "n_comp": self.n_comp}

    def load_state_dict(self, f):
        self.model.load_state_dict(f)

    # ---- reference helpers kept for API compatibility (small torch ops, not on the hot path) ----
    def compute_beta_m(self, U, V, b, withbias=True, tonp=False):
        if tonp:
            U, V = np2tensor(U), np2tensor(V)
        beta = U @ V
        if withbias:
            if tonp:
                b = np2tensor(b)
        else:
            b = torch.zeros((U.shape[0], 1, V.shape[1]), dtype=beta.dtype, device=beta.device)
        beta = torch.cat((beta, b), 1)
        return tensor2np(beta) if tonp else beta

    def compute_beta(self, eid, withbias=True):
        return self.compute_beta_m(self.model[f"{eid}_U"], self.model['V'], self.model[f"{eid}_b"], withbias=withbias)

    def predict(self, beta, X, tonp=False):
        if tonp:
            X, beta = np2tensor(X), np2tensor(beta)
        y_pred = torch.einsum("ktc,nct->ktn", X, beta)
        return tensor2np(y_pred) if tonp else y_pred

    # ---- device-resident data --------------------------------------------------------------
    def _device(self):
        dev = self.model['V'].device
        if dev.type != "cuda":
            raise vs.VsError("RRRGD parameters are not on a CUDA device: call .to(get_device()) first (no CPU path)")
        return dev

    def _split(self, data, eid, k):
        X = data[eid]['X'][k]
        if isinstance(X, _PackedSplit):
            return X
        key = (eid, k, id(X))
        hit = self._packed.get(key)
        if hit is None:
            hit = _PackedSplit(X, data[eid]['y'][k], self.n_comp, self.planes, self._device(), self.fmt)
            self._packed[key] = hit
        return hit

    def _workspace(self, dims):
        need = int(vs.lib.vs_rrr_workspace(dims))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self._device())
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:off + need]

    def _closure_eval(self, data, eid, k, want_grad, dV=None):
        """loss (0-dim), sse_n (N,) and -- if want_grad -- dU, db written into .grad, dV accumulated."""
        sp = self._split(data, eid, k)
        sp.wait_ready()
        if sp.overflow is not None and self.n_closure_evals > 0:
            sp.check_range()
        U, b, V = self.model[f"{eid}_U"], self.model[f"{eid}_b"], self.model['V']
        dev = V.device
        loss = torch.empty(1, dtype=torch.float64, device=dev)
        sse = torch.empty(sp.N, dtype=torch.float64, device=dev)
        dU = db = None
        if want_grad:
            dU, db = self._grad_buffer(U), self._grad_buffer(b)
        if sp.dims.mode != vs.RRR_MODE_CLASSIC:
            import ctypes
            groups = sp.neuron_groups()
            lossg = loss if len(groups) == 1 else torch.empty(len(groups), dtype=torch.float64, device=dev)
            for gi, (n0, n1, dg, yg, ylog) in enumerate(groups):
                ws = self._workspace(dg)
                ops = sp.exact_ops(ylog)
                sl = lambda t_: vs.ptr(t_[n0:n1]) if t_ is not None else None      # noqa: E731  (contiguous neuron slices)
                vs.check(vs.lib.vs_rrr_closure_exact(dg, vs.ptr(sp.Xa), ctypes.byref(ops), vs.ptr(sp.xl), vs.ptr(yg), sl(U.data),
                                                     vs.ptr(V.data), sl(b.data), float(self.l2), vs.ptr(lossg[gi:gi + 1]), sl(sse), sl(dU),
                                                     vs.ptr(dV), sl(db), vs.ptr(ws), ws.numel(), vs.stream()))
            if len(groups) > 1:
                # every group's scalar carries the V-only part of nothing (the l2 term is <U_g^T U_g, V V^T> + |b_g|^2): plain sum
                loss = lossg.sum().reshape(1)
            return loss[0], sse, dU, db
        ws = self._workspace(sp.dims)
        vs.check(vs.lib.vs_rrr_closure(sp.dims, vs.ptr(sp.Xa), vs.ptr(sp.Xb), vs.ptr(sp.xl), vs.ptr(sp.y), vs.ptr(U.data),
                                       vs.ptr(V.data), vs.ptr(b.data), float(self.l2), vs.ptr(loss), vs.ptr(sse), vs.ptr(dU),
                                       vs.ptr(dV), vs.ptr(db), self.engine, vs.ptr(ws), ws.numel(), vs.stream()))
        return loss[0], sse, dU, db

    @staticmethod
    def _grad_buffer(p):
        """`p.grad` if it can be written in place (FusedLBFGS keeps views of its flat gradient there), else a
        fresh tensor installed as `p.grad`.  The kernels overwrite every element."""
        g = p.grad
        if g is None or not g.is_contiguous() or g.dtype != p.dtype or g.device != p.device or g.shape != p.shape:
            g = torch.empty_like(p.data)
            p.grad = g
        return g

    def loss_and_grad(self, data, k=0):
        """The closure body of rrr.py:165-175: total loss over sessions, gradients into .grad."""
        V = self.model['V']
        dV = self._grad_buffer(V)
        dV.zero_()                                  # accumulated over the sessions sharing V (rrr.py:46-49)
        total = None
        for eid in data:
            loss, _, dU, db = self._closure_eval(data, eid, k, True, dV)
            total = loss if total is None else total + loss
        for eid in self.eids:                       # sessions the caller left out contribute nothing: their gradients are zero,
            if eid not in data:                     # never stale (make_optimizer() promises that every element is written)
                for name in (f"{eid}_U", f"{eid}_b"):
                    self._grad_buffer(self.model[name]).zero_()
        self.n_closure_evals += 1
        return total

    # ---- reference API ------------------------------------------------------------------------
    def predict_y(self, data, eid, k):
        """rrr.py:122-130.  Returns (X, y, ypred); y / ypred are float64 CUDA tensors.  X is returned
        as the caller's own array wrapped in a CPU tensor (the reference copies it to the device on
        every call; no caller in the reference uses that copy)."""
        sp = self._split(data, eid, k)
        sp.wait_ready()
        U, b, V = self.model[f"{eid}_U"], self.model[f"{eid}_b"], self.model['V']
        yhat = torch.empty((sp.K, sp.T, sp.N), dtype=torch.float64, device=V.device)
        groups = sp.neuron_groups() if sp.dims.mode != vs.RRR_MODE_CLASSIC else None
        if groups is not None and len(groups) > 1:
            for n0, n1, dg, _, _ in groups:
                ws = self._workspace(dg)
                yg = torch.empty((sp.K, sp.T, n1 - n0), dtype=torch.float64, device=V.device)
                vs.check(vs.lib.vs_rrr_predict(dg, vs.ptr(sp.Xa), vs.ptr(sp.xl), vs.ptr(U.data[n0:n1]), vs.ptr(V.data), vs.ptr(b.data[n0:n1]),
                                               vs.ptr(yg), self.engine, vs.ptr(ws), ws.numel(), vs.stream()))
                yhat[:, :, n0:n1] = yg
            ws = None
        elif sp.dims.mode == vs.RRR_MODE_DENSE:
            ws = self._workspace(sp.dims)
            import ctypes
            ops = sp.exact_ops()
            vs.check(vs.lib.vs_rrr_predict_exact(sp.dims, ctypes.byref(ops), vs.ptr(sp.xl), vs.ptr(U.data), vs.ptr(V.data), vs.ptr(b.data),
                                                 vs.ptr(yhat), vs.ptr(ws), ws.numel(), vs.stream()))
        else:
            ws = self._workspace(sp.dims)
            vs.check(vs.lib.vs_rrr_predict(sp.dims, vs.ptr(sp.Xa), vs.ptr(sp.xl), vs.ptr(U.data), vs.ptr(V.data), vs.ptr(b.data),
                                           vs.ptr(yhat), self.engine, vs.ptr(ws), ws.numel(), vs.stream()))
        Xraw = data[eid]['X'][k]
        X = np2tensor(Xraw) if isinstance(Xraw, np.ndarray) else None
        y = np2tensor(np.ascontiguousarray(data[eid]['y'][k])).to(V.device) if isinstance(data[eid]['y'][k], np.ndarray) \
            else sp.y.double()
        return X, y, yhat

    def predict_y_fr(self, data, eid, k):
        """rrr.py:136-142: back to firing-rate units with the stored z-score statistics."""
        X, y, ypred = self.predict_y(data, eid, k)
        mean_y = torch.as_tensor(data[eid]['setup']['mean_y_TN']).to(y.device)
        std_y = torch.as_tensor(data[eid]['setup']['std_y_TN']).to(y.device)
        return X, y * std_y + mean_y, ypred * std_y + mean_y

    def compute_MSE_RRRGD(self, data, k):
        """rrr.py:147-152: {eid: per-neuron sum of squared residuals (N,)}."""
        return {eid: self._closure_eval(data, eid, k, False)[1] for eid in data}

    def regression_loss(self):
        """rrr.py:154-155: {eid: l2 * sum(beta^2)} -- evaluated from the r x r Gram matrices, beta is
        never materialised: sum (U V)^2 = <U^T U, V V^T>."""
        V = self.model['V']
        W = V @ V.T
        out = {}
        for eid in self.eids:
            U, b = self.model[f"{eid}_U"], self.model[f"{eid}_b"]
            G = torch.einsum("nci,ncj->ij", U, U)
            out[eid] = self.l2 * (torch.sum(G * W) + torch.sum(b ** 2))
        return out


def train_model(model, train_data, optimizer, model_fname, save=True):
    """rrr.py:164-190: one optimizer.step(closure) on split 0, validation SSE on split 1."""
    def closure():
        optimizer.zero_grad()
        model.train()
        return model.loss_and_grad(train_data, 0)

    optimizer.step(closure)

    model.eval()
    mses_val = model.compute_MSE_RRRGD(train_data, 1)
    best_loss = torch.sum(torch.cat([mses_val[k] for k in mses_val]))

    if save:
        print('saving model')
        torch.save({"RRRGD_model": model.state_dict(), "optimizer": optimizer.state_dict()}, model_fname)

    return model, {"mses_val": mses_val, "mse_val_mean": best_loss}


def train_model_main(train_data, l2, n_comp, model_fname, save=True, planes=None, engine=None, operand=None):
    """rrr.py:192-202 (the parameters are created on the device directly; `.to(device)` is then a no-op)."""
    device = get_device()
    area_model = RRRGD(train_data, n_comp, l2=l2, planes=planes, engine=engine, operand=operand, device=device)
    area_model.to(device)
    print(f"training on device: {device}")
    optimizer = area_model.make_optimizer()                    # torch.optim.LBFGS semantics, device-side vector algebra
    _, mse_val = train_model(area_model, train_data, optimizer, model_fname=model_fname, save=save)
    return area_model, mse_val


# ------------------------------------------------------------------------------------------------
# R0 on the device: the preprocessing of src/train_rrr.py:108-171 for the video modalities, from raw
# uint8 frames, without ever forming the float64 (K, T, C) matrix on the host.
def pack_session_from_frames(frames_train, counts_train, frames_test, counts_test, sorted_idx, n_comp, planes=None,
                             smooth_w=2.0, device=None, operand=None, mode=None, stats_hook=None):
    """frames_*: uint8 (K, Tf, ...) torch tensors (pinned host or CUDA); counts_*: (K, T, N) spike counts.
    Mirrors train_rrr.py: y smoothed with gaussian_filter1d(sigma=smooth_w, axis=1); X and y z-scored with
    the TRAIN statistics (std clipped at 1e-8); ones column; frames `sorted_idx` selected AFTER the z-score.
    Returns the per-session entry of the `train_data` dict RRRGD consumes, with device-resident splits.
    `mode` (see rrr_mode): "exact" unless `planes` is given (then "classic" with that many planes).
    `stats_hook(mean, sd, my, sy, K) -> (mean, sd, my, sy)`: called once with the TRAIN statistics (frames and smoothed
    counts, float64, std clipped) before anything is z-scored; parallel.pack_trial_shard combines the ranks' statistics there."""
    vs.require_b200()
    if mode is None:
        mode = "classic" if planes is not None else rrr_mode(None)
    mode = rrr_mode(mode)
    exact = mode in ("exact", "dense")
    dense = mode == "dense"
    if exact:
        # sessions wider than one launch of the exact-operand kernels (160 neurons) are evaluated in neuron groups
        # (_PackedSplit.neuron_groups); shapes outside those kernels altogether (rank != 3, <= 128 features): the parity mode
        # is then the classic layout with 3 residual planes
        Kq, Fq, Nq = int(frames_train.shape[0]), int(frames_train[0, 0].numel()), int(counts_train.shape[2])
        bounds = _neuron_group_bounds(Nq)
        if len(bounds) > 1:
            dense = False                       # neuron groups are evaluated by the factorised exact closure
        if not vs.lib.vs_rrr_exact_supported(Kq, len(np.asarray(sorted_idx)), Fq, max(b1 - b0 for b0, b1 in bounds), n_comp):
            exact, planes = False, 3
    if exact:
        planes, operand = 2, "f16"
    elif planes is None:
        planes = int(os.environ.get("VS_RRR_PLANES", "1"))
    device = device or torch.device("cuda")
    st = vs.stream()
    idx_np = np.ascontiguousarray(np.asarray(sorted_idx), dtype=np.int32)
    idx = torch.as_tensor(idx_np).to(device)
    T = int(idx.numel())
    idx_compact = torch.arange(T, dtype=torch.int32, device=device)
    host_select = T > 0 and bool(np.all(np.diff(idx_np) > 0))       # vs_h2d_select_frames wants strictly increasing indices
    compact_stats = False
    fmt = operand_format(planes, operand)
    splits, ys = [], []
    mean = sd = my = sy = None
    main = torch.cuda.current_stream(device)
    side = _side_stream(device)
    for which, (fr, cnt) in enumerate(((frames_train, counts_train), (frames_test, counts_test))):
        # the fit needs the TRAIN split only: the test split (H2D + pack) goes to a side stream that starts once the train
        # statistics exist, overlaps the parameter upload and the first closure evaluations, and is joined lazily
        # (_PackedSplit.wait_ready) by the first kernel that reads it
        K, Tf = int(fr.shape[0]), int(fr.shape[1])
        F = int(fr[0, 0].numel())
        N = int(cnt.shape[2])
        d = vs.RrrDims(K, T, F, N, n_comp, planes, vs.lib.vs_rrr_ldc(F), vs.lib.vs_rrr_ldr(K, T), fmt,
                       (vs.RRR_MODE_DENSE if dense else vs.RRR_MODE_EXACT) if exact else vs.RRR_MODE_CLASSIC)
        ex = None
        if exact:
            # forward operand: z as hi + lo half planes (exact) or the exact integers in the same row layout (dense);
            # backward operand: ONE plane of exact integers, train split only (the other splits are only evaluated);
            # the scale tables come from the train statistics and are shared by every split of the session
            Xa = None if dense else torch.empty((planes, K * T, d.ldc), dtype=_op_dtype(fmt), device=device)
            Xb = torch.empty((F, d.ldr), dtype=_op_dtype(fmt), device=device) if which == 0 else None
            ldt = int(vs.lib.vs_rrr_ldt(T))
            if which == 0:
                tables = {"isdT": torch.empty((F, ldt), dtype=torch.float32, device=device),
                          "qT": torch.empty((F, ldt), dtype=torch.float32, device=device)}
                if dense:
                    Tq = (T + 15) // 16 * 16
                    tables.update({"isd": torch.empty((T, d.ldc), dtype=torch.float32, device=device),
                                   "qh": torch.empty((2, Tq, d.ldc), dtype=torch.float16, device=device),
                                   "isdmax": torch.empty(T, dtype=torch.float32, device=device)})
            ex = dict(tables)
            ex.update({"ldt": ldt, "y_lo": torch.empty((K, T, N), dtype=torch.float32, device=device)})
            if dense:
                ex["Xc"] = torch.empty((K * T, d.ldc), dtype=_op_dtype(fmt), device=device)
        else:
            Xa = torch.empty((planes, K * T, d.ldc), dtype=_op_dtype(fmt), device=device)          # allocated on the main stream
            Xb = torch.empty((planes, F, d.ldr), dtype=_op_dtype(fmt), device=device)
        xl = torch.empty(K * T, dtype=torch.float32, device=device)
        y = torch.empty((K, T, N), dtype=torch.float32, device=device)
        overflow = torch.zeros(1, dtype=torch.int32, device=device)
        if which == 1:
            side.wait_stream(main)                  # (buffers are main-stream allocations: _PackedSplit.__del__ joins the
            #                                         side stream if the split dies unread, so they are never recycled early)
        with torch.cuda.stream(side if which == 1 else main):
            st = vs.stream()
            if host_select and not fr.is_cuda and fr.dtype == torch.uint8 and fr.is_contiguous() and T < Tf \
                    and (which == 0 or compact_stats):
                # only the T selected frames of each trial cross PCIe (one strided DMA per run of consecutive indices); the
                # z-score statistics of a frame depend on that frame alone, so the unselected ones are never needed
                frd = torch.empty((K, T, F), dtype=torch.uint8, device=device)
                vs.check(vs.lib.vs_h2d_select_frames(fr.data_ptr(), K, Tf, F, idx_np.ctypes.data, T, vs.ptr(frd), st))
                fr, Tf_dev, idx_dev = frd, T, idx_compact
            else:
                fr = fr.to(device, non_blocking=True).reshape(K, Tf, -1).contiguous()
                Tf_dev, idx_dev = Tf, idx
                if which == 1 and compact_stats:              # train statistics are compact: bring the test frames to the same order
                    fr = fr[:, idx.long()].contiguous()
                    Tf_dev, idx_dev = T, idx_compact
            cnt = torch.as_tensor(cnt).to(device, non_blocking=True).float().contiguous()
            # fused loader (one read of the frames: statistics + pack per time bin) when every frame on the device is a
            # selected one; otherwise the two-kernel path (statistics of all Tf frames, then the pack)
            fused = (exact and Tf_dev == T and F % 4 == 0 and K * 136 <= 200 * 1024 and fr.data_ptr() % 4 == 0
                     and os.environ.get("VS_RRR_FUSED_PACK", "1") != "0")
            if which == 0:
                compact_stats = Tf_dev == T and Tf != T
                mean = torch.empty(Tf_dev * F, dtype=torch.float64, device=device)
                sd = torch.empty_like(mean)
                if not fused or stats_hook is not None:
                    vs.check(vs.lib.vs_rrr_colstats(vs.ptr(fr), K, Tf_dev * F, vs.ptr(mean), vs.ptr(sd), st))
                sm = torch.empty_like(cnt)
                vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, N, float(smooth_w), None, None, vs.ptr(sm), st))
                my = torch.empty(T * N, dtype=torch.float64, device=device)
                sy = torch.empty_like(my)
                vs.check(vs.lib.vs_colstats_f32(vs.ptr(sm), K, T * N, vs.ptr(my), vs.ptr(sy), st))
                del sm
                if stats_hook is not None:
                    mean, sd, my, sy = (t.contiguous() for t in stats_hook(mean, sd, my, sy, K))
            if exact:
                import ctypes
                tw = (lambda k: vs.ptr(ex[k]) if (which == 0 and k in ex) else None)      # tables are written with the train split only
                out = vs.RrrExactOps(vs.ptr(Xb) if Xb is not None else None, tw("isdT"), tw("qT"), ldt, None,
                                     vs.ptr(ex["Xc"]) if dense else None, tw("isd"), tw("qh"), tw("isdmax"))
                if fused:
                    vs.check(vs.lib.vs_rrr_pack_u8_fused(vs.ptr(fr), Tf_dev, vs.ptr(idx_dev), vs.ptr(mean), vs.ptr(sd),
                                                         int(which == 0 and stats_hook is None), d,
                                                         vs.ptr(Xa) if Xa is not None else None, ctypes.byref(out), vs.ptr(xl),
                                                         vs.ptr(overflow), st))
                else:
                    vs.check(vs.lib.vs_rrr_pack_u8_exact(vs.ptr(fr), Tf_dev, vs.ptr(idx_dev), vs.ptr(mean), vs.ptr(sd), d,
                                                         vs.ptr(Xa) if Xa is not None else None, ctypes.byref(out), vs.ptr(xl),
                                                         vs.ptr(overflow), st))
                vs.check(vs.lib.vs_rrr_smooth_y2(vs.ptr(cnt), K, T, N, float(smooth_w), vs.ptr(my), vs.ptr(sy), vs.ptr(y),
                                                 vs.ptr(ex["y_lo"]), st))
            else:
                vs.check(vs.lib.vs_rrr_pack_u8(vs.ptr(fr), Tf_dev, vs.ptr(idx_dev), vs.ptr(mean), vs.ptr(sd), d, vs.ptr(Xa), vs.ptr(Xb),
                                               vs.ptr(xl), vs.ptr(overflow), st))
                vs.check(vs.lib.vs_rrr_smooth_y(vs.ptr(cnt), K, T, N, float(smooth_w), vs.ptr(my), vs.ptr(sy), vs.ptr(y), st))
            if exact and len(_neuron_group_bounds(N)) > 1:
                ex["groups"] = []
                for n0, n1 in _neuron_group_bounds(N):
                    dg = vs.RrrDims(K, T, F, n1 - n0, n_comp, planes, d.ldc, d.ldr, fmt, d.mode)
                    ex["groups"].append((n0, n1, dg, y[:, :, n0:n1].contiguous(), ex["y_lo"][:, :, n0:n1].contiguous()))
            sp = _PackedSplit.from_device(d, Xa, Xb, xl, y, overflow, exact=ex)
            if which == 1:
                sp._keep = (idx, idx_compact, idx_dev, mean, sd, my, sy)
                sp.ready = torch.cuda.Event()
                sp.ready.record(side)
            del fr, cnt
        splits.append(sp)
        ys.append(_LazyY(sp) if which == 1 else (y.double() + ex["y_lo"].double() if exact else y))
    if compact_stats:
        # API shape of the reference's `setup` (per frame of the trial window): statistics exist for the selected frames only
        Tf_all, F_all = int(frames_train.shape[1]), int(frames_train[0, 0].numel())
        mean_full = torch.full((Tf_all, F_all), float("nan"), dtype=torch.float64, device=device)
        sd_full = torch.full((Tf_all, F_all), float("nan"), dtype=torch.float64, device=device)
        mean_full[idx.long()] = mean.view(T, F_all)
        sd_full[idx.long()] = sd.view(T, F_all)
        mean, sd = mean_full.reshape(-1), sd_full.reshape(-1)
    return {"X": splits, "y": ys,
            "setup": {"mean_X_Tv": mean, "std_X_Tv": sd, "mean_y_TN": my.reshape(T, -1), "std_y_TN": sy.reshape(T, -1)}}


def train_model_from_frames(frames_train, counts_train, frames_test, counts_test, sorted_idx, l2=100.0, n_comp=3,
                            eid="session", planes=None, engine=None, model_fname="tmp", save=False, operand=None, mode=None):
    """R0 + train_model_main (rrr.py:192-202) in one call, from raw uint8 frames (`mode`, `planes`: pack_session_from_frames)."""
    entry = pack_session_from_frames(frames_train, counts_train, frames_test, counts_test, sorted_idx, n_comp, planes=planes,
                                     operand=operand, mode=mode)
    train_data = {eid: entry}
    model, mse_val = train_model_main(train_data, l2, n_comp, model_fname, save=save, planes=planes, engine=engine, operand=operand)
    return model, mse_val, train_data
