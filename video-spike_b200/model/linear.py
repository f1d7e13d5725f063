"""`Linear` video->spike model: drop-in for the reference's src/model/linear.py:3-56.

Same constructor (`Linear(config.model)`), same sub-module tree, so `state_dict()` keys
(`encoder.layers.{0,2,4}.{weight,bias}`, `decoder.layers.{0,2,4}.{weight,bias}`), parameter
order and default `nn.Linear` initialisation under a given torch seed are identical to the
reference.  What differs is where the arithmetic runs: forward, backward and (with
`optim.FusedAdamW`) the update are the sm_100a kernels of libvs_b200.so.  There is no CPU path.

Input: uint8 frames (B, T, 1, H, W) straight from the loader (the `.float()` of
src/loader/base.py:39 is folded into the kernels), or float32 of any (B, ...) shape as the
reference trainer passes it (src/trainer/base.py:64-67).
"""
from __future__ import annotations

import ctypes as C

import torch

import vsb200 as vs


def _mlp_stack(config):
    """src/model/linear.py:17-36 / 38-56: Linear+ReLU per hidden dim, then a bare Linear."""
    layer_num = config.layer_num
    hidden_dims = config.hidden_dims
    assert len(hidden_dims) == layer_num, "hidden_dims must have the same length as layer_num"
    layers = torch.nn.Sequential()
    prev = config.input_dim
    for width in hidden_dims:
        layers.append(torch.nn.Linear(prev, width))
        layers.append(torch.nn.ReLU())
        prev = width
    layers.append(torch.nn.Linear(prev, config.output_dim))
    return layers


class Encoder(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = _mlp_stack(config)

    def forward(self, x):
        raise RuntimeError("Encoder is evaluated as part of Linear.forward (fused MLP kernels)")


class Decoder(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = _mlp_stack(config)

    def forward(self, x):
        raise RuntimeError("Decoder is evaluated as part of Linear.forward (fused MLP kernels)")


def _layer_list(model):
    """[(nn.Linear, relu_after)] over encoder then decoder, in forward order."""
    out = []
    for seq in (model.encoder.layers, model.decoder.layers):
        mods = list(seq)
        for i, m in enumerate(mods):
            if isinstance(m, torch.nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
                out.append((m, relu))
    return out


class _MlpBuffers:
    """Activation / gradient scratch for one batch size."""

    def __init__(self, layers, batch, device, train):
        n = len(layers)
        self.batch = batch
        self.act = [torch.empty((batch, lin.out_features), dtype=torch.float32, device=device) for lin, _ in layers]
        self.gact = [torch.empty_like(a) for a in self.act] if train else [None] * n
        # gradients are only materialised on the large-batch route (batch > 32); at the reference batch
        # size they live in registers inside the fused dW+AdamW kernels
        big = train and batch > 32
        self.gW = [torch.empty_like(lin.weight) if big else None for lin, _ in layers]
        self.gb = [torch.empty_like(lin.bias) if (big and lin.bias is not None) else None for lin, _ in layers]
        self.loss = torch.zeros(1, dtype=torch.float64, device=device)


class _MlpFunction(torch.autograd.Function):
    """Autograd bridge used when the model is driven by stock torch code (reference trainer,
    torch.optim.AdamW): forward = vs_mlp_forward, backward = vs_linear_bwd per layer."""

    @staticmethod
    def forward(ctx, model, x, *params):
        logits, acts = model._forward_impl(x)
        ctx.model, ctx.x, ctx.acts = model, x, acts
        return logits

    @staticmethod
    def backward(ctx, gout):
        model, x, acts = ctx.model, ctx.x, ctx.acts
        layers = model._layers
        g = gout.contiguous().float().reshape(gout.shape[0], -1)
        grads = [None] * (2 * len(layers))
        st = vs.stream()
        for l in range(len(layers) - 1, -1, -1):
            lin, relu = layers[l]
            xin = x if l == 0 else acts[l - 1]
            x_f32 = xin if xin.dtype == torch.float32 else None
            x_u8 = xin if xin.dtype == torch.uint8 else None
            batch = g.shape[0]
            gm = torch.empty_like(g) if relu else g
            dx = torch.empty((batch, lin.in_features), dtype=torch.float32, device=g.device) if l > 0 else None
            defer = (l == 0 and getattr(lin.weight, "_vs_defer_ok", False) and batch <= 32
                     and lin.in_features % 4 == 0 and lin.out_features * 128 <= 200 * 1024)
            dW = None if defer else torch.empty_like(lin.weight)
            db = torch.empty_like(lin.bias) if lin.bias is not None else None
            need = int(vs.lib.vs_linear_bwd_workspace(batch, lin.in_features, lin.out_features)) if dx is not None else 0
            ws = torch.empty(need, dtype=torch.uint8, device=g.device) if need else None
            vs.check(vs.lib.vs_linear_bwd(vs.ptr(g), vs.ptr(acts[l]), vs.ptr(x_f32), vs.ptr(x_u8), vs.ptr(lin.weight),
                                          vs.ptr(gm), vs.ptr(dx), vs.ptr(dW), vs.ptr(db), batch, lin.in_features,
                                          lin.out_features, int(relu), vs.ptr(ws), need, st))
            if defer:
                # the first-layer gradient stays factored as (dy, x); FusedAdamW consumes it without
                # ever materialising the (256, D) matrix
                lin.weight._vs_lowrank_grad = (gm, xin)
            grads[2 * l] = dW
            grads[2 * l + 1] = db
            g = dx
        return (None, None, *grads)


class Linear(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder = Encoder(config.encoder)
        self.decoder = Decoder(config.decoder)
        self.output_dim = config.decoder.output_dim // 100   # 100 time bins hard-coded (linear.py:8)
        self.engine = vs.ENGINE_AUTO

    # checkpoints pickle the whole module (src/trainer/base.py:285-291): drop device scratch
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_train_bufs", "_layers_cache", "_ws"):
            state.pop(k, None)
        return state

    @property
    def _layers(self):
        cache = self.__dict__.get("_layers_cache")
        if cache is None:
            cache = _layer_list(self)
            self.__dict__["_layers_cache"] = cache
        return cache

    def _net(self, bufs, opt_state=None):
        layers = self._layers
        net = vs.Mlp()
        net.n_layers = len(layers)
        net.dims[0] = layers[0][0].in_features
        for l, (lin, relu) in enumerate(layers):
            net.dims[l + 1] = lin.out_features
            net.relu[l] = int(relu)
            net.W[l] = vs.ptr(lin.weight)
            net.b[l] = vs.ptr(lin.bias) if lin.bias is not None else None
            net.act[l] = vs.ptr(bufs.act[l])
            net.gact[l] = vs.ptr(bufs.gact[l])
            net.gW[l] = vs.ptr(bufs.gW[l])
            net.gb[l] = vs.ptr(bufs.gb[l])
            if opt_state is not None:
                sw = opt_state[lin.weight]
                net.mW[l], net.vW[l] = vs.ptr(sw["exp_avg"]), vs.ptr(sw["exp_avg_sq"])
                if lin.bias is not None:
                    sb = opt_state[lin.bias]
                    net.mb[l], net.vb[l] = vs.ptr(sb["exp_avg"]), vs.ptr(sb["exp_avg_sq"])
        return net

    @staticmethod
    def _split_input(x):
        if not x.is_cuda:
            raise vs.VsError("video-spike_b200 Linear runs on CUDA tensors only (no CPU fallback)")
        x = x.flatten(1).contiguous()                          # linear.py:11
        if x.dtype == torch.uint8:
            return None, x
        return x.float(), None

    def _workspace(self, net, batch, device):
        need = int(vs.lib.vs_mlp_workspace(C.byref(net), batch))
        ws = self.__dict__.get("_ws")
        if ws is None or ws.numel() < need or ws.device != device:
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device=device)
            self.__dict__["_ws"] = ws
        return ws

    def _forward_impl(self, x, target=None):
        """vs_mlp_forward into freshly allocated activations (autograd / eval keep them alive)."""
        x_f32, x_u8 = self._split_input(x)
        batch, device = x.shape[0], x.device
        bufs = _MlpBuffers(self._layers, batch, device, False)
        net = self._net(bufs)
        ws = self._workspace(net, batch, device)
        vs.check(vs.lib.vs_mlp_forward(C.byref(net), vs.ptr(x_u8), vs.ptr(x_f32), vs.ptr(target), batch,
                                       vs.ptr(bufs.loss) if target is not None else None, self.engine, vs.ptr(ws),
                                       ws.numel(), vs.stream()))
        logits = bufs.act[-1].reshape(-1, 100, self.output_dim)   # linear.py:14
        return logits, bufs.act

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = []
            for lin, _ in self._layers:
                params += [lin.weight, lin.bias]
            xin = x.flatten(1).contiguous()
            if xin.dtype != torch.uint8:
                xin = xin.float()
            return _MlpFunction.apply(self, xin, *params)
        return self._forward_impl(x)[0]

    # ---- row-parallel first layer ---------------------------------------------------------------
    def shard_first_layer(self, rank, world):
        """Keep only this rank's pixel slice of the first layer: W0[:, lo:hi] (the Adam moments follow when the
        optimizer is built afterwards).  Slices are multiples of 1024 pixels so every rank's update kernel sees whole
        column strips.  Everything after the first pre-activation stays replicated.  Returns (lo, hi)."""
        lin = self.encoder.layers[0]
        D = lin.in_features
        chunk = -(-D // (world * 1024)) * 1024
        lo, hi = min(D, rank * chunk), min(D, (rank + 1) * chunk)
        if hi <= lo:
            raise vs.VsError(f"rank {rank} of {world} gets no pixels of a {D}-pixel frame")
        self.__dict__["_shard"] = {"lo": lo, "hi": hi, "D": D, "rank": rank, "world": world}
        if lo == 0 and hi == D:
            return lo, hi                      # the whole layer: keep the Parameter (an existing optimizer stays valid)
        with torch.no_grad():
            w = lin.weight[:, lo:hi].clone().contiguous()
        lin.weight = torch.nn.Parameter(w)
        lin.in_features = hi - lo
        self.__dict__["_shard"] = {"lo": lo, "hi": hi, "D": D, "rank": rank, "world": world}
        for k in ("_train_bufs", "_layers_cache", "_ws"):
            self.__dict__.pop(k, None)
        return lo, hi

    def fused_train_step_rowpar(self, x, target, optimizer, group=None, allreduce=None):
        """One training step with the first layer row-parallel: phase 0 (local partial pre-activation), ONE
        all-reduce(sum) of the (B, 256) fp32 pre-activation, phase 1 (everything else, all local).  `x` is this
        rank's pixel slice (B, hi-lo) or the whole frames (B, D) (sliced here).  `allreduce(tensor)` overrides
        torch.distributed.all_reduce (used by the single-GPU emulation test)."""
        sh = self.__dict__.get("_shard")
        if sh is None:
            raise vs.VsError("call shard_first_layer(rank, world) first")
        x = x.flatten(1)
        if x.shape[1] == sh["D"] and sh["D"] != sh["hi"] - sh["lo"]:
            x = x[:, sh["lo"]:sh["hi"]]
        x_f32, x_u8 = self._split_input(x)
        batch, device = x.shape[0], x.device
        cache = self.__dict__.setdefault("_train_bufs", {})
        bufs = cache.get((batch, str(device)))
        if bufs is None:
            bufs = cache[(batch, str(device))] = _MlpBuffers(self._layers, batch, device, True)
        hyper = optimizer.begin_fused_step()
        net = self._net(bufs, optimizer.state)
        ws = self._workspace(net, batch, device)
        target = target.contiguous().float()
        args = (C.byref(net), vs.ptr(x_u8), vs.ptr(x_f32), vs.ptr(target), batch, hyper, vs.ptr(bufs.loss), self.engine, vs.ptr(ws),
                ws.numel(), vs.stream())
        vs.check(vs.lib.vs_mlp_train_step_rowpar(*args, 0))
        if allreduce is not None:
            allreduce(bufs.act[0])
        elif sh["world"] > 1:
            torch.distributed.all_reduce(bufs.act[0], group=group)
        vs.check(vs.lib.vs_mlp_train_step_rowpar(*args, 1))
        return bufs.loss[0] / float(target.numel())

    def fused_train_step(self, x, target, optimizer):
        """One whole step of src/trainer/base.py:147-154 (forward, Poisson NLL, backward, AdamW on
        every parameter) as ONE C-ABI call (vs_mlp_train_step).  Returns the mean loss as a
        0-dim float64 CUDA tensor (no host sync)."""
        x_f32, x_u8 = self._split_input(x)
        batch, device = x.shape[0], x.device
        cache = self.__dict__.setdefault("_train_bufs", {})
        bufs = cache.get((batch, str(device)))
        if bufs is None:
            bufs = cache[(batch, str(device))] = _MlpBuffers(self._layers, batch, device, True)
        hyper = optimizer.begin_fused_step()
        net = self._net(bufs, optimizer.state)
        ws = self._workspace(net, batch, device)
        target = target.contiguous().float()
        vs.check(vs.lib.vs_mlp_train_step(C.byref(net), vs.ptr(x_u8), vs.ptr(x_f32), vs.ptr(target), batch, hyper,
                                          vs.ptr(bufs.loss), self.engine, vs.ptr(ws), ws.numel(), vs.stream()))
        return bufs.loss[0] / float(target.numel())
