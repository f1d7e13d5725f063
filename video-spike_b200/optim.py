"""FusedAdamW: torch.optim.AdamW semantics (src/train.py:44-49) executed by libvs_b200 kernels.

It is a regular `torch.optim.Optimizer` (same `param_groups` keys: lr, betas, eps, weight_decay),
so `OneCycleLR` can drive it exactly as in the reference -- including the beta1 cycling that
`cycle_momentum=True` applies to Adam-family optimizers: lr and beta1 are read from the
group at every step and passed to the kernels as per-step scalars.

Two ways to consume gradients:
  * `step()`                -- after a normal autograd backward.  A first layer whose gradient was left
                               factored by model.linear._MlpFunction (`weight._vs_lowrank_grad`) is
                               updated by vs_dw_adamw_fused (dW never materialised); every other
                               parameter by vs_adamw.
  * `begin_fused_step()`    -- used by Linear.fused_train_step: bumps the step counters and returns
                               the hyper-parameters; the update itself happens inside vs_mlp_train_step.
"""
from __future__ import annotations

import ctypes as C

import torch

import vsb200 as vs


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        for group in self.param_groups:
            for p in group["params"]:
                p._vs_defer_ok = True   # lets the first layer keep its gradient factored

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            if not p.is_cuda:
                raise vs.VsError("FusedAdamW only updates CUDA parameters (no CPU fallback)")
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @staticmethod
    def _hyper(group, step):
        b1, b2 = group["betas"]
        return vs.AdamWHyper(float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                             float(group["weight_decay"]), int(step))

    def begin_fused_step(self):
        if len(self.param_groups) != 1:
            raise vs.VsError("fused train step supports a single parameter group (as src/train.py builds)")
        group = self.param_groups[0]
        step = None
        for p in group["params"]:
            st = self._state_for(p)
            st["step"] += 1
            step = st["step"]
        return self._hyper(group, step)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stream = vs.stream()
        for group in self.param_groups:
            for p in group["params"]:
                low = getattr(p, "_vs_lowrank_grad", None)
                if p.grad is None and low is None:
                    continue
                st = self._state_for(p)
                st["step"] += 1
                h = self._hyper(group, st["step"])
                if low is not None:
                    dy, x = low
                    x_f32 = x if x.dtype == torch.float32 else None
                    x_u8 = x if x.dtype == torch.uint8 else None
                    vs.check(vs.lib.vs_dw_adamw_fused(vs.ptr(dy), vs.ptr(x_f32), vs.ptr(x_u8), vs.ptr(p.data),
                                                      vs.ptr(st["exp_avg"]), vs.ptr(st["exp_avg_sq"]), None, None, None, dy.shape[0],
                                                      p.shape[1], p.shape[0], h, stream))
                    p._vs_lowrank_grad = None
                else:
                    g = p.grad.contiguous()
                    vs.check(vs.lib.vs_adamw(vs.ptr(p.data), vs.ptr(g), vs.ptr(st["exp_avg"]), vs.ptr(st["exp_avg_sq"]),
                                             p.numel(), h, stream))
        return loss

    def zero_grad(self, set_to_none: bool = True):
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_vs_lowrank_grad", None) is not None:
                    p._vs_lowrank_grad = None
        super().zero_grad(set_to_none=set_to_none)


# =====================================================================================================
# L-BFGS (R5)
def lbfgs_two_loop(gg, sg, yg, SY, YY, H_diag):
    """torch.optim.LBFGS's two-loop recursion in COEFFICIENT space (host, float64, m <= 100).

    The direction d = -H*g is a linear combination of the basis {g, s_0..s_{m-1}, y_0..y_{m-1}}; the
    recursion only needs inner products of basis vectors:  sg[i] = s_i.g, yg[i] = y_i.g,
    SY[i][j] = s_i.y_j, YY[i][j] = y_i.y_j, gg = g.g.  Returns (cg, cs[m], cy[m], gtd) with
    d = cg*g + sum cs[i]*s_i + cy[i]*y_i and gtd = g.d.  Pure numpy: unit-tested on the CPU against
    the explicit vector recursion."""
    import numpy as np
    m = len(sg)
    ro = [1.0 / SY[i][i] for i in range(m)]
    al = [0.0] * m
    # q = -g - sum_j al_j y_j  (coefficients: g -> -1, y_j -> -al_j)
    for i in range(m - 1, -1, -1):
        sq = -sg[i]
        for j in range(i + 1, m):
            sq -= al[j] * SY[i][j]
        al[i] = sq * ro[i]
    # r = H*q + sum_j (al_j - be_j) s_j
    cs = [0.0] * m
    for i in range(m):
        yr = -yg[i]
        for j in range(m):
            yr -= al[j] * YY[i][j]
        yr *= H_diag
        for j in range(i):
            yr += cs[j] * SY[j][i]
        cs[i] = al[i] - yr * ro[i]
    cg = -H_diag
    cy = [-H_diag * a for a in al]
    gtd = cg * gg + float(np.dot(cs, sg)) + float(np.dot(cy, yg)) if m else cg * gg
    return cg, cs, cy, gtd


class FusedLBFGS(torch.optim.Optimizer):
    """torch.optim.LBFGS (no line search) as src/model/rrr.py:177,199 uses it -- same constructor, same
    `step(closure)` contract, same iteration/termination logic -- with the vector algebra on libvs_b200:

      * all parameters live in ONE flat float64 buffer (each `p.data` is rebound to a view of it) and the
        gradient in a second one (`p.grad` views): no gather/scatter per iteration.  Two gradient buffers
        ping-pong so "prev_flat_grad" is never copied.
      * per iteration the device makes two streaming passes (vs_lbfgs_dots, vs_lbfgs_direction); the
        two-loop recursion runs on the host in coefficient space (lbfgs_two_loop).
      * one host<->device synchronisation per iteration (the scalars of vs_lbfgs_dots + the loss).

    `history_dtype=torch.float32` stores the curvature pairs (s_i, y_i) in float32 (all arithmetic stays float64):
    half the HBM traffic of both passes; the default float64 reproduces torch.optim.LBFGS to rounding.

    `device_driven=True` moves the decisions themselves (memory update, two-loop recursion, step length, termination
    tests) into a one-thread kernel working on a device-resident state (vs_lbfgs_dev_*): a whole `step(closure)` is then
    ENQUEUED without any host<->device synchronisation -- the host only runs ahead calling the closure -- and the state is
    read back once at the end.  After the state says "done" the remaining launches are no-ops, so an early-terminated
    step still calls the closure max_iter times (on unchanged parameters); `state["func_evals"]` / `["n_iter"]` report
    the logical counts.  Same arithmetic, same order: trajectories agree with the host-driven mode to rounding.

    A closure that wants to avoid allocations writes gradients INTO the existing `p.grad` views
    (model.rrr.RRRGD.loss_and_grad does); autograd closures work too (`zero_grad()` zeroes the views)."""

    def __init__(self, params, lr=1, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9,
                 history_size=100, line_search_fn=None, history_dtype=torch.float64, device_driven=False, compact=False):
        """`compact=True` (device-driven, float64 history): ONE stored vector per closure evaluation -- the gradient
        differences -- instead of the (s_i, y_i) pairs; the steps live as coefficient rows and every inner product of
        the recursion comes from the basis Gram matrix (csrc/lbfgs.cu "compact history").  Same decisions in the same
        order, about half the HBM traffic of the vector passes; at most vs.LBFGS_CMAX closure evaluations over the life
        of the optimiser (the reference runs one step of <= 25)."""
        if history_dtype not in (torch.float64, torch.float32):
            raise ValueError("history_dtype must be torch.float64 or torch.float32")
        self._hdtype = history_dtype
        self._device_driven = bool(device_driven)
        self._compact = bool(compact)
        if self._compact and (not self._device_driven or history_dtype != torch.float64 or history_size < vs.LBFGS_CMAX):
            raise vs.VsError("FusedLBFGS(compact=True) needs device_driven=True, a float64 history and history_size >= %d" % vs.LBFGS_CMAX)
        self._dev = None
        if line_search_fn is not None:
            raise vs.VsError("FusedLBFGS implements the fixed-step variant only (the reference never sets line_search_fn)")
        if history_size > 100:
            raise vs.VsError("FusedLBFGS: history_size > 100 (VS_LBFGS_MAX_HIST) is not supported")
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        super().__init__(params, dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                                      tolerance_change=tolerance_change, history_size=history_size,
                                      line_search_fn=line_search_fn))
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]["params"]
        self._flat = None
        self._pairs = []          # [(s_slot, y_slot)] oldest first
        self._new_gram()          # _SY[i, j] = s_i . y_j, _YY[i, j] = y_i . y_j for the pairs in _pairs
        self._hist = None
        self._free = []

    def _new_gram(self):
        import numpy as np
        h = self.param_groups[0]["history_size"]
        self._SY = np.zeros((h, h))
        self._YY = np.zeros((h, h))
        self._coef = np.zeros(2 * h + 1)

    def _two_loop(self, m, gg, sg, yg, H_diag):
        """-> (coef[0 .. 2m], gtd): the compiled copy of lbfgs_two_loop (csrc/host_lbfgs.cpp)."""
        import numpy as np
        sg = np.ascontiguousarray(sg, dtype=np.float64)
        yg = np.ascontiguousarray(yg, dtype=np.float64)
        gtd = C.c_double()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)                                  # noqa: E731
        vs.check(vs.lib.vs_host_lbfgs_two_loop(m, float(gg), vp(sg), vp(yg), vp(self._SY), vp(self._YY), self._SY.shape[1],
                                               float(H_diag), vp(self._coef), C.byref(gtd)))
        return self._coef[:2 * m + 1], gtd.value

    # ---- flat storage -----------------------------------------------------------------------------
    def _bind(self):
        ps = self._params
        dev = ps[0].device
        for p in ps:
            self._check_param(p, dev)
        fl = self._flat
        if fl is not None and all(p.data_ptr() == fl["x"].data_ptr() + 8 * a for p, (a, _) in zip(ps, fl["span"])):
            return fl
        n = sum(p.numel() for p in ps)
        x = torch.empty(n, dtype=torch.float64, device=dev)
        g = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(2)]
        span, a = [], 0
        for p in ps:
            b = a + p.numel()
            x[a:b].copy_(p.data.reshape(-1))
            if p.grad is not None:
                g[0][a:b].copy_(p.grad.reshape(-1))
            p.data = x[a:b].view(p.shape)
            span.append((a, b))
            a = b
        self._flat = fl = {"x": x, "g": g, "cur": 0, "span": span, "n": n}
        self._point_grads(0)
        m = self.param_groups[0]["history_size"]
        self._scal = torch.zeros(8 + 6 * m + 8, dtype=torch.float64, device=dev)      # dots out | dmax | loss
        self._alloc_workspace(n, m, dev)
        self._pairs, self._hist, self._free = [], None, []
        self._dev = None               # the device-resident state (slot tables, history window) belonged to the old storage
        self._new_gram()
        self.state[ps[0]].clear()
        return fl

    def _point_grads(self, which):
        fl = self._flat
        fl["cur"] = which
        for p, (a, b) in zip(self._params, fl["span"]):
            p.grad = fl["g"][which][a:b].view(p.shape)

    def zero_grad(self, set_to_none: bool = True):
        """Keeps the `p.grad` views alive (a reference-style closure calls this first, rrr.py:166).  A closure whose kernels
        OVERWRITE every gradient element (RRRGD.loss_and_grad) declares so with `closure_overwrites_grads = True`; the
        63 MB memset per evaluation is then skipped."""
        if self._flat is None:
            return super().zero_grad(set_to_none=set_to_none)
        if getattr(self, "closure_overwrites_grads", False):
            return
        self._flat["g"][self._flat["cur"]].zero_()

    def _slot(self):
        """A free history slot; the buffer grows in pairs-of-8 steps (each slot is one flat vector)."""
        if not self._free:
            n = self._flat["n"]
            old = self._hist
            have = 0 if old is None else old.shape[0]
            grow = 2 * min(self.param_groups[0]["history_size"] + 1, max(8, self.param_groups[0]["max_iter"] + 1))
            npad = (n + 3) // 4 * 4                                     # pitch of 4 elements: every slot 16-byte aligned
            new = torch.empty((have + grow, npad), dtype=self._hdtype, device=self._flat["x"].device)
            if old is not None:
                new[:have].copy_(old)
            self._hist = new
            self._free = list(range(have, have + grow))
        return self._free.pop(0)

    # ---- vector passes (libvs_b200; the hooks exist so parallel.ShardedLBFGS can restrict / all-reduce them and
    #      the CPU tests can drive the host logic with a torch backend) -----------------------------------------
    def _check_param(self, p, dev):
        if not p.is_cuda or p.dtype != torch.float64 or p.device != dev:
            raise vs.VsError("FusedLBFGS needs float64 CUDA parameters on one device (no CPU path)")

    def _alloc_workspace(self, n, m, dev):
        self._ws = torch.empty(int(vs.lib.vs_lbfgs_workspace(n, m)), dtype=torch.uint8, device=dev)

    def _slot_arrays(self):
        m = len(self._pairs)
        ss = (C.c_int32 * max(m, 1))(*[p[0] for p in self._pairs])
        ys = (C.c_int32 * max(m, 1))(*[p[1] for p in self._pairs])
        return m, ss, ys

    def _pass_dots(self, lo, hi, g, g_prev, s_slot, y_slot, out):
        """vs_lbfgs_dots over elements [lo, hi) of the flat vectors: y_slot <- g - g_prev, scalars -> out."""
        m, ss, ys = self._slot_arrays()
        hist = self._hist
        esz = 4 if self._hdtype == torch.float32 else 8
        off = lambda t, e: None if t is None else t.data_ptr() + lo * e        # noqa: E731
        for t in (g, g_prev, hist):
            if t is not None:
                vs.ptr(t)                                                        # CUDA + contiguity check
        vs.check(vs.lib.vs_lbfgs_dots(hi - lo, off(g, 8), off(g_prev, 8), off(hist[s_slot], esz) if s_slot is not None else None,
                                      off(hist[y_slot], esz) if y_slot is not None else None, off(hist, esz),
                                      hist.shape[1] if hist is not None else hi - lo, int(self._hdtype == torch.float32), ss, ys, m,
                                      vs.ptr(out), vs.ptr(self._ws), self._ws.numel(), vs.stream()))

    def _pass_direction(self, g, coef, t, x, s_slot, dmax_out):
        """vs_lbfgs_direction over the whole flat vector: d = sum coef*basis; hist[s_slot] <- t*d; x += t*d."""
        m, ss, ys = self._slot_arrays()
        hist = self._hist
        import numpy as np
        cf = np.ascontiguousarray(coef, dtype=np.float64).ctypes.data_as(C.c_void_p)
        vs.check(vs.lib.vs_lbfgs_direction(self._flat["n"], vs.ptr(g), vs.ptr(hist), hist.shape[1], int(self._hdtype == torch.float32),
                                           ss, ys, m, cf, float(t), vs.ptr(x) if x is not None else None, vs.ptr(hist[s_slot]),
                                           vs.ptr(dmax_out), vs.stream()))

    def _reduce_scalars(self, k):
        """Hook: combine self._scal[:k] (sums; index 2 is a max) and self._scal[-2] (max) across ranks."""

    def _dots(self, g, g_prev, s_slot, y_slot, loss_t):
        """-> host numpy array [dots out (8+6m) ..., dmax, loss] after ONE synchronisation."""
        k = 8 + 6 * len(self._pairs)
        self._pass_dots(0, self._flat["n"], g, g_prev, s_slot, y_slot, self._scal)
        self._reduce_scalars(k)
        self._scal[-1:].copy_(torch.as_tensor(loss_t).detach().reshape(1))
        host = self._scal.cpu().numpy()
        return host[:k], float(host[-2]), float(host[-1])

    # ---- device-driven step ----------------------------------------------------------------------------------
    _DEV_HEADER = 1728        # leading bytes of vs_lbfgs_dev: counters, flags, slot tables, scalars (up to and including dmax)

    def _dev_state(self, need_slots):
        """Device buffer holding a vs_lbfgs_dev (+ float64 views of its `out` block and `dmax`), with a history buffer
        of at least `need_slots` slots.  Grows between steps: new slot ids are appended to the state's free list."""
        dev_t = self._flat["x"].device
        n = self._flat["n"]
        npad = (n + 3) // 4 * 4
        need_slots = min(need_slots, 2 * vs.LBFGS_MAX_HIST + 8)
        if self._dev is None:
            host = vs.LbfgsDev()
            vs.check(vs.lib.vs_lbfgs_dev_init_host(C.byref(host), need_slots))
            raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(dev_t)
            k = 8 + 6 * vs.LBFGS_MAX_HIST
            self._dev = {"raw": raw, "n_slots": need_slots, "host": host,
                         "out": raw[vs.LbfgsDev.out.offset:vs.LbfgsDev.out.offset + 8 * k].view(torch.float64),
                         "dmax": raw[vs.LbfgsDev.dmax.offset:vs.LbfgsDev.dmax.offset + 8].view(torch.float64),
                         "ws": torch.empty(int(vs.lib.vs_lbfgs_dev_workspace(n)), dtype=torch.uint8, device=dev_t)}
            self._hist = torch.empty((need_slots, npad), dtype=self._hdtype, device=dev_t)
        elif need_slots > self._dev["n_slots"]:
            old_n, host = self._dev["n_slots"], self._dev["host"]
            grown = torch.empty((need_slots, npad), dtype=self._hdtype, device=dev_t)
            grown[:old_n].copy_(self._hist)
            self._hist = grown
            for s_id in range(old_n, need_slots):
                host.free_slots[host.n_free] = s_id
                host.n_free += 1
            head = torch.frombuffer(bytearray(bytes(host)[:self._DEV_HEADER]), dtype=torch.uint8)
            self._dev["raw"][:self._DEV_HEADER].copy_(head.to(dev_t))
            self._dev["n_slots"] = need_slots
        return self._dev

    def _read_dev_state(self):
        head = self._dev["raw"][:self._DEV_HEADER].cpu().numpy().tobytes()
        host = vs.LbfgsDev.from_buffer_copy(head.ljust(C.sizeof(vs.LbfgsDev), b"\0"))
        self._dev["host"] = host
        return host

    def _reduce_dev_scalars(self):
        """Hook (parallel.ShardedLBFGS): combine the state's `out` block and `dmax` across ranks on the device."""

    def _dots_call(self, sp, n, g_ptr, gp_ptr, hist_ptr, stride, f32, dev):
        """One dots pass over `n` elements starting at the given device addresses (classic or compact state)."""
        if dev.get("compact"):
            vs.check(vs.lib.vs_lbfgs_cdev_dots(sp, n, g_ptr, gp_ptr, hist_ptr, stride, vs.ptr(dev["ws"]), dev["ws"].numel(), vs.stream()))
        else:
            vs.check(vs.lib.vs_lbfgs_dev_dots(sp, n, g_ptr, gp_ptr, hist_ptr, stride, f32, vs.ptr(dev["ws"]), dev["ws"].numel(), vs.stream()))

    def _dev_dots(self, sp, n, g, g_prev, hist, f32, dev):
        """The dots pass over the whole flat vector (parallel.ShardedLBFGS restricts the inner products of the
        replicated prefix to rank 0)."""
        self._dots_call(sp, n, vs.ptr(g), vs.ptr(g_prev), vs.ptr(hist), hist.shape[1], f32, dev)

    def _cdev_state(self, need_slots):
        """Device buffer holding a vs_lbfgs_cdev (+ float64 views of its `out` block and `dmax`) and the basis vectors."""
        dev_t = self._flat["x"].device
        n = self._flat["n"]
        npad = (n + 3) // 4 * 4
        need_slots = min(need_slots, vs.LBFGS_CMAX)
        if self._dev is None:
            host = vs.LbfgsCDev()
            vs.check(vs.lib.vs_lbfgs_cdev_init_host(C.byref(host)))
            raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(dev_t)
            k = 8 + 3 * vs.LBFGS_CMAX
            self._dev = {"raw": raw, "n_slots": need_slots, "host": host, "compact": True,
                         "out": raw[vs.LbfgsCDev.out.offset:vs.LbfgsCDev.out.offset + 8 * k].view(torch.float64),
                         "dmax": raw[vs.LbfgsCDev.dmax.offset:vs.LbfgsCDev.dmax.offset + 8].view(torch.float64),
                         "ws": torch.empty(int(vs.lib.vs_lbfgs_cdev_workspace(n)), dtype=torch.uint8, device=dev_t)}
            self._hist = torch.empty((need_slots, npad), dtype=torch.float64, device=dev_t)
        elif need_slots > self._dev["n_slots"]:
            grown = torch.empty((need_slots, npad), dtype=torch.float64, device=dev_t)
            grown[:self._dev["n_slots"]].copy_(self._hist)
            self._hist = grown
            self._dev["n_slots"] = need_slots
        return self._dev

    @torch.no_grad()
    def _step_compact(self, closure):
        closure = torch.enable_grad()(closure)
        group = self.param_groups[0]
        lr, max_iter, max_eval = float(group["lr"]), group["max_iter"], group["max_eval"]
        tol_g, tol_c = float(group["tolerance_grad"]), float(group["tolerance_change"])
        fl = self._bind()
        n = fl["n"]
        state = self.state[self._params[0]]
        nb_now = self._dev["host"].nb if self._dev is not None else 0
        if nb_now + max_iter > vs.LBFGS_CMAX:
            raise vs.VsError("FusedLBFGS(compact=True): more than %d closure evaluations over the life of the optimiser; "
                             "use compact=False" % vs.LBFGS_CMAX)
        dev = self._cdev_state(nb_now + max_iter + 1)
        hist = self._hist
        sp, st = vs.ptr(dev["raw"]), vs.stream
        orig_loss = None
        head = vs.LbfgsCDev.coef.offset
        for it in range(max_iter):
            prev = state.get("prev_buf")
            if prev is not None and fl["cur"] == prev:
                self._point_grads(1 - prev)
            loss_t = closure()
            if orig_loss is None:
                orig_loss = loss_t
            g = fl["g"][fl["cur"]]
            g_prev = fl["g"][1 - fl["cur"]]
            lt = torch.as_tensor(loss_t).detach().reshape(1).double()
            self._dev_dots(sp, n, g, g_prev, hist, 0, dev)
            self._reduce_dev_scalars()
            vs.check(vs.lib.vs_lbfgs_cdev_update(sp, vs.ptr(lt), lr, tol_g, tol_c, int(max_eval), int(it == 0), st()))
            vs.check(vs.lib.vs_lbfgs_cdev_direction(sp, n, vs.ptr(g), vs.ptr(hist), hist.shape[1], vs.ptr(fl["x"]), st()))
            state["prev_buf"] = fl["cur"]
        raw = dev["raw"][:head].cpu().numpy().tobytes()               # the only synchronisation of the step
        host = vs.LbfgsCDev.from_buffer_copy(raw.ljust(C.sizeof(vs.LbfgsCDev), b"\0"))
        dev["host"] = host
        if host.done == 7:
            raise vs.VsError("FusedLBFGS(compact=True): basis full (vs.LBFGS_CMAX evaluations)")
        state["func_evals"], state["n_iter"] = int(host.func_evals), int(host.total_iter)
        state["dev_done"] = int(host.done)
        return orig_loss

    @torch.no_grad()
    def _step_device_driven(self, closure):
        closure = torch.enable_grad()(closure)
        group = self.param_groups[0]
        lr, max_iter, max_eval = float(group["lr"]), group["max_iter"], group["max_eval"]
        tol_g, tol_c, hsize = float(group["tolerance_grad"]), float(group["tolerance_change"]), group["history_size"]
        fl = self._bind()
        n = fl["n"]
        state = self.state[self._params[0]]
        m_now = self._dev["host"].m if self._dev is not None else 0
        dev = self._dev_state(2 * min(hsize, m_now + max_iter) + 6)
        hist, f32 = self._hist, int(self._hdtype == torch.float32)
        sp, st = vs.ptr(dev["raw"]), vs.stream
        orig_loss = None
        for it in range(max_iter):
            # the closure writes the new gradient into the buffer that does not hold the previous one
            prev = state.get("prev_buf")
            if prev is not None and fl["cur"] == prev:
                self._point_grads(1 - prev)
            loss_t = closure()
            if orig_loss is None:
                orig_loss = loss_t
            g = fl["g"][fl["cur"]]
            g_prev = fl["g"][1 - fl["cur"]]
            lt = torch.as_tensor(loss_t).detach().reshape(1).double()
            self._dev_dots(sp, n, g, g_prev, hist, f32, dev)
            self._reduce_dev_scalars()
            vs.check(vs.lib.vs_lbfgs_dev_update(sp, vs.ptr(lt), lr, tol_g, tol_c, int(max_eval), int(hsize), int(it == 0), st()))
            vs.check(vs.lib.vs_lbfgs_dev_direction(sp, n, vs.ptr(g), vs.ptr(hist), hist.shape[1], f32, vs.ptr(fl["x"]), st()))
            state["prev_buf"] = fl["cur"]
        host = self._read_dev_state()                        # the only synchronisation of the step
        state["func_evals"], state["n_iter"] = int(host.func_evals), int(host.total_iter)
        state["dev_done"] = int(host.done)
        return orig_loss

    @torch.no_grad()
    def step(self, closure):
        if self._device_driven and self._compact:
            return self._step_compact(closure)
        if self._device_driven:
            return self._step_device_driven(closure)
        import numpy as np
        closure = torch.enable_grad()(closure)
        group = self.param_groups[0]
        lr, max_iter, max_eval = float(group["lr"]), group["max_iter"], group["max_eval"]
        tol_g, tol_c, hsize = group["tolerance_grad"], group["tolerance_change"], group["history_size"]
        fl = self._bind()
        state = self.state[self._params[0]]
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)
        n = fl["n"]

        def evaluate():
            """closure at the current x, written into the gradient buffer that does NOT hold torch's
            `prev_flat_grad` (state["prev_buf"]); returns the scalars of one vs_lbfgs_dots pass."""
            prev = state.get("prev_buf")
            if prev is not None and fl["cur"] == prev:
                self._point_grads(1 - prev)
            loss_t = closure()
            g = fl["g"][fl["cur"]]
            g_prev = fl["g"][prev] if prev is not None else None
            s_slot = state.get("s_slot")
            y_slot = self._slot() if (g_prev is not None and s_slot is not None) else None
            out, dmax, loss = self._dots(g, g_prev, s_slot if y_slot is not None else None, y_slot, loss_t)
            return loss_t, loss, out, dmax, y_slot

        orig_loss, loss, out, _, y_slot = evaluate()
        current_evals = 1
        state["func_evals"] += 1
        if out[2] <= tol_g:
            if y_slot is not None:
                self._free.append(y_slot)
            return orig_loss
        H_diag = state.get("H_diag", 1.0)
        prev_loss = state.get("prev_loss")
        n_iter = 0
        while n_iter < max_iter:
            n_iter += 1
            state["n_iter"] += 1
            m = len(self._pairs)
            gg, g1 = float(out[0]), float(out[1])
            hs = out[8:8 + 6 * m].reshape(2 * m, 3)        # rows: s_0..s_{m-1}, y_0..y_{m-1}; cols: .g, .y_new, .s_new
            sg, yg = hs[:m, 0].copy(), hs[m:, 0].copy()
            if state["n_iter"] == 1:
                H_diag = 1.0
            elif y_slot is not None:
                yy, ys_new = float(out[3]), float(out[4])
                s_slot = state["s_slot"]
                if ys_new > 1e-10:
                    sy_col, yy_col, ys_row = hs[:m, 1], hs[m:, 1], hs[m:, 2]           # s_i.y_new, y_i.y_new, y_i.s_new
                    lo = 0
                    if m == hsize:                                                     # limited memory: drop the oldest
                        self._free.extend(self._pairs.pop(0))
                        self._SY[:m - 1, :m - 1] = self._SY[1:m, 1:m].copy()
                        self._YY[:m - 1, :m - 1] = self._YY[1:m, 1:m].copy()
                        lo, m = 1, m - 1
                    self._SY[:m, m] = sy_col[lo:]
                    self._SY[m, :m] = ys_row[lo:]
                    self._SY[m, m] = ys_new
                    self._YY[:m, m] = yy_col[lo:]
                    self._YY[m, :m] = yy_col[lo:]
                    self._YY[m, m] = yy
                    self._pairs.append((s_slot, y_slot))
                    sg = np.append(sg[lo:], out[5])
                    yg = np.append(yg[lo:], out[6])
                    H_diag = ys_new / yy
                    state["s_slot"] = None
                else:
                    self._free.extend([s_slot, y_slot])
                    state["s_slot"] = None
                y_slot = None
            m = len(self._pairs)
            coef, gtd = self._two_loop(m, gg, sg, yg, H_diag)
            state["prev_buf"] = fl["cur"]          # torch: prev_flat_grad.copy_(flat_grad) -- here a buffer swap
            prev_loss = loss
            t = min(1.0, 1.0 / g1) * lr if state["n_iter"] == 1 else lr
            g = fl["g"][fl["cur"]]
            if state.get("s_slot") is not None:
                self._free.append(state["s_slot"])
            s_slot = self._slot()
            stop_gtd = gtd > -tol_c
            self._pass_direction(g, coef, t, None if stop_gtd else fl["x"], s_slot, self._scal[-2:-1])
            state["s_slot"] = s_slot
            if stop_gtd:
                break
            ls_func_evals = 0
            if n_iter != max_iter:
                _, loss, out, dmax, y_slot = evaluate()
                ls_func_evals = 1
            current_evals += ls_func_evals
            state["func_evals"] += ls_func_evals
            if n_iter == max_iter:
                break
            if current_evals >= max_eval:
                break
            if out[2] <= tol_g:
                break
            if dmax <= tol_c:
                break
            if abs(loss - prev_loss) < tol_c:
                break
        # a y vector computed by the last evaluation but not consumed: recomputed by the next step()'s first pass
        if y_slot is not None:
            self._free.append(y_slot)
        state["H_diag"] = H_diag
        state["prev_loss"] = prev_loss
        return orig_loss
