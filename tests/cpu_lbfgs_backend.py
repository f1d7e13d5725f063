"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the two device passes of FusedLBFGS (vs_lbfgs_dots /
vs_lbfgs_direction), with the same output layout, so the HOST logic of the optimiser -- Gram bookkeeping,
coefficient-space recursion, termination, buffer ping-pong, and ShardedLBFGS's all-reduces over gloo -- can be
exercised without a GPU.  Never imported by the product path."""
import torch


class TorchPasses:
    def _check_param(self, p, dev):
        assert p.dtype == torch.float64

    def _alloc_workspace(self, n, m, dev):
        self._ws = None

    def _pass_dots(self, lo, hi, g, g_prev, s_slot, y_slot, out):
        hist = self._hist
        gv = g[lo:hi]
        pv = g_prev[lo:hi] if g_prev is not None else torch.zeros_like(gv)
        yv = (gv - pv).to(self._hdtype).double()
        sv = hist[s_slot][lo:hi].double() if s_slot is not None else torch.zeros_like(gv)
        if y_slot is not None:
            hist[y_slot][lo:hi] = yv.to(self._hdtype)
        m = len(self._pairs)
        res = [gv @ gv, gv.abs().sum(), gv.abs().max() if gv.numel() else torch.zeros(()), yv @ yv, yv @ sv, sv @ gv, yv @ gv, torch.zeros(())]
        for which in (0, 1):
            for i in range(m):
                h = hist[self._pairs[i][which]][lo:hi].double()
                res += [h @ gv, h @ yv, h @ sv]
        out[:8 + 6 * m] = torch.stack([r.double().reshape(()) for r in res])

    def _pass_direction(self, g, coef, t, x, s_slot, dmax_out):
        n = self._flat["n"]
        m = len(self._pairs)
        d = coef[0] * g[:n]
        for i in range(m):
            d = d + coef[1 + i] * self._hist[self._pairs[i][0]][:n].double() + coef[1 + m + i] * self._hist[self._pairs[i][1]][:n].double()
        sd = t * d
        self._hist[s_slot][:n] = sd.to(self._hdtype)
        if x is not None:
            x += sd
        dmax_out[0] = sd.abs().max()
