"""Whole-network forward/backward schedules for the ST-CGAN generator (8-level U-Net) and the
70x70 PatchGAN discriminator, expressed as sequences of the C-ABI kernels in `ops`.

What the reference does with nested nn.Sequential + autograd (src/models/stcgan_g.py:55-57,120-132;
src/models/stcgan_d.py:57-58) is here an explicit schedule over NHWC buffers:

  * skip concatenations are never produced by a copy: the encoder's BN/activation pass writes
    relu(x_k) straight into the left half of the decoder's input buffer, the decoder's BN pass
    writes relu(bn(u_{k+1})) into the right half (cropped, for odd sizes);
  * the in-place LeakyReLU / ReLU pair on the skip becomes two outputs of one pass;
  * odd spatial sizes are handled by letting the convolution read its zero padding out of
    range (no F.pad copy) and by cropping inside the BN/activation pass.

Nothing here computes with torch; torch only owns the buffers.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, BACKEND_FFMA, BACKEND_TC, GEOM_PARITY,
                   GEOM_WIN_S1, GEOM_WIN_S1_FLIP, GEOM_WIN_S2)

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class ConvOp:
    """One Conv2d(4,s,1) / ConvTranspose2d(4,2,1) weight with its packed copies and packed gradient."""

    def __init__(self, kind, weight, bias):
        assert kind in ("conv2", "conv1", "convT")
        self.kind, self.weight, self.bias = kind, weight, bias
        self.d0, self.d1 = int(weight.shape[0]), int(weight.shape[1])
        self.cin, self.cout = (self.d0, self.d1) if kind == "convT" else (self.d1, self.d0)
        self.p1 = self.p2 = None          # packed weights [16,d0,d1], [16,d1,d0]
        self.g = None                     # packed gradient [16,d0,d1] fp32 (view into the net's flat buffer)
        self.gb = None                    # bias gradient
        self.version = None
        self.use_tc = False
        # thin layers on the tensor cores (bf16): "cin" = Conv2d with <= 8 input channels (first layers),
        # "coutT" = ConvTranspose2d with <= 8 output channels (G's last layer), "cout1" = stride-1 Conv2d with <= 8
        # output channels (D's last layer).  See include/stcgan_b200.h "thin layers".
        self.thin = None
        self.wthin = self.wtn = None
        self.cpad = 0

    # ---- packing ------------------------------------------------------------------------
    def ensure_packed(self, act_dtype, force=False):
        w = self.weight
        ver = (w._version, w.data_ptr(), act_dtype)
        if force or ver != self.version or self.p1 is None:
            if self.p1 is None or self.p1.dtype != act_dtype or self.p1.device != w.device:
                self.p1 = torch.empty((16, self.d0, self.d1), dtype=act_dtype, device=w.device)
                self.p2 = torch.empty((16, self.d1, self.d0), dtype=act_dtype, device=w.device)
            wc = w.detach() if w.is_contiguous() else w.detach().contiguous()
            ops.pack_weight(wc, self.p1, self.p2)
            self.thin = None
            if act_dtype == torch.bfloat16:
                if self.kind == "conv2" and self.cin <= 8 and self.cout % 64 == 0:
                    self.thin = "cin"
                elif self.kind == "convT" and self.cout <= 8 and self.cin % 64 == 0:
                    self.thin = "coutT"
                elif self.kind == "conv1" and self.cout <= 8 and self.cin % 64 == 0:
                    self.thin = "cout1"
            if self.thin is not None:
                fat = self.cout if self.thin == "cin" else self.cin
                nthin = self.cin if self.thin == "cin" else self.cout
                self.cpad = 1 if nthin == 1 else (4 if nthin <= 4 else 8)
                if self.wthin is None or self.wthin.device != w.device:
                    self.wthin = torch.empty((fat, 128), dtype=torch.bfloat16, device=w.device)
                    self.wtn = torch.empty((16 * self.cpad, fat), dtype=torch.bfloat16, device=w.device)
                # wthin: thin-K operand [fat][(tap, c)];  wtn: thin-N operand [(tap, c)][fat] of the pixel GEMM + col2im path
                if self.thin == "cin":        # fwd: thin K over ci;  dgrad: thin N over ci (= d1)
                    ops.pack_weight_thin(wc, True, False, self.wthin)
                    ops.pack_weight_tapn(wc, False, self.cpad, self.wtn)
                elif self.thin == "coutT":    # fwd: thin N over co (= d1);  dgrad: thin K over co, rows ci (= d0)
                    ops.pack_weight_tapn(wc, False, self.cpad, self.wtn)
                    ops.pack_weight_thin(wc, True, False, self.wthin)
                else:                         # fwd: thin N over co (= d0);  dgrad: thin K over co with flipped taps
                    ops.pack_weight_tapn(wc, True, self.cpad, self.wtn)
                    ops.pack_weight_thin(wc, False, True, self.wthin)
            self.version = ver
        self.use_tc = act_dtype == torch.bfloat16

    def mark_packed(self):
        """The optimiser kernel refreshed p1/p2 from the updated weight: record the new version as packed."""
        w = self.weight
        if self.version is not None and self.thin is None:
            self.version = (w._version, w.data_ptr(), self.version[2])

    def _backend_conv(self, k, nout, plain_epilogue=True):
        return BACKEND_TC if (self.use_tc and plain_epilogue and ops.tc_eligible_conv(k, nout)) else BACKEND_FFMA

    # ---- the three GEMMs ------------------------------------------------------------------
    def out_size(self, ih, iw):
        if self.kind == "conv2":
            return (ih + 2 - 4) // 2 + 1, (iw + 2 - 4) // 2 + 1
        if self.kind == "conv1":
            return ih - 1, iw - 1
        return ih * 2, iw * 2

    def stats_fusable(self):
        """True if forward() can accumulate the following BatchNorm's batch statistics in its GEMM epilogue."""
        return self.use_tc and self.thin is None and self.bias is None and ops.tc_eligible_conv(self.cin, self.cout)

    def forward(self, x, oh, ow, *, out=None, out_nchw=None, act=ACT_NONE, use_bias=True, x_bordered=None, bn_acc=None):
        geom = {"conv2": GEOM_WIN_S2, "conv1": GEOM_WIN_S1, "convT": GEOM_PARITY}[self.kind]
        wp = self.p2 if self.kind == "convT" else self.p1
        bias = self.bias.detach() if (self.bias is not None and use_bias) else None
        if self.thin == "cin" and x_bordered is not None and out_nchw is None:
            return ops.thinconv(x_bordered, 2, self.wthin, self.cout, oh, ow, bias=bias, act=act, out=out)
        if self.thin in ("coutT", "cout1") and out_nchw is not None:
            ops.thin_col2im(0 if self.thin == "coutT" else 1, x, self.wtn, self.cpad, self.cout, oh, ow, bias=bias, act=act,
                            out_nchw=out_nchw)
            return out_nchw
        plain = out_nchw is None and act in (ACT_NONE, ACT_LEAKY, ACT_RELU)
        return ops.tapconv(geom, x, wp, self.cout, oh, ow, bias=bias, act=act, out=out, out_nchw=out_nchw,
                           backend=self._backend_conv(self.cin, self.cout, plain), bn_acc=bn_acc)

    def dgrad(self, g, ih, iw, *, out=None, out8=None, g_bordered=None):
        """gradient w.r.t. the layer input [N, ih, iw, cin] from g = gradient w.r.t. the layer output.
        Thin layers: `out8` = full 8-channel NHWC tensor for the thin-cin case; `g_bordered` = zero-bordered 8-channel
        output gradient for the thin-cout cases."""
        geom = {"conv2": GEOM_PARITY, "conv1": GEOM_WIN_S1_FLIP, "convT": GEOM_WIN_S2}[self.kind]
        wp = self.p1 if self.kind == "convT" else self.p2
        if self.thin == "cin" and out8 is not None:
            ops.thin_col2im(0, g, self.wtn, self.cpad, self.cin, ih, iw, out8=out8)
            return out8
        if self.thin in ("coutT", "cout1") and g_bordered is not None:
            return ops.thinconv(g_bordered, 2 if self.thin == "coutT" else 1, self.wthin, self.cin, ih, iw, out=out)
        return ops.tapconv(geom, g, wp, self.cin, ih, iw, out=out, backend=self._backend_conv(self.cout, self.cin))

    def wgrad(self, x, g, *, x_bordered=None, g_bordered=None, bias_done=False):
        """accumulate the packed weight gradient from the layer input x and output gradient g (`bias_done`: the bias
        gradient was already accumulated by the activation-backward pass that produced g)."""
        if self.thin == "cin" and x_bordered is not None:
            ops.thinwgrad(x_bordered, 2, self.cin, g, self.g, True, False)
            if self.bias is not None and self.gb is not None and not bias_done:
                ops.colsum(g, self.gb)
            return
        if self.thin in ("coutT", "cout1") and g_bordered is not None:
            ops.thinwgrad(g_bordered, 2 if self.thin == "coutT" else 1, self.cout, x, self.g,
                          self.thin == "coutT", self.thin == "cout1")
            if self.bias is not None and self.gb is not None:
                ops.colsum(g_bordered[..., :self.cout], self.gb)
            return
        if self.kind == "convT":
            s, l, geom = x, g, GEOM_WIN_S2
        else:
            s, l, geom = g, x, (GEOM_WIN_S2 if self.kind == "conv2" else GEOM_WIN_S1)
        backend = BACKEND_TC if (self.use_tc and ops.tc_eligible_wgrad(self.d0, self.d1)) else BACKEND_FFMA
        ops.tapwgrad(geom, s, l, self.g, backend=backend)
        if self.bias is not None and self.gb is not None:
            ops.colsum(g, self.gb)


class BNOp:
    """BatchNorm2d parameters/buffers + per-pass scratch handling."""

    def __init__(self, bn_module):
        self.m = bn_module
        self.c = bn_module.num_features
        self.ggamma = self.gbeta = None     # fp32 gradient views
        self.shared_counter = False         # True: num_batches_tracked is a view into the runtime's flat counter tensor

    def forward(self, y, scratch, training, out1, act1, out2=None, act2=ACT_NONE, stats_done=False, deferred=None):
        """[stats (training, unless the producing GEMM already accumulated them: stats_done)] -> finalize + apply in one
        launch.  scratch: dict with 'acc' (f64 [BN_SLOTS,2,C], zeroed), 'mi', 'ss' (f32 [2,C]).
        `deferred` (a list): the running-statistics update is NOT performed; (self, acc, count) is appended instead and
        `apply_running_update` performs it later (the caller orders it after a concurrently running pass)."""
        m = self.m
        n, h, w, _ = y.shape
        if training:
            if n * h * w <= 1:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {[n, self.c, h, w]}")
            if not stats_done:
                ops.bn_stats(y, scratch["acc"])          # fills slot 0
            use_running = m.track_running_stats and m.running_mean is not None
            now = use_running and deferred is None
            ops.bn_fused_apply(y, scratch["acc"], n * h * w, m.weight.detach(), m.bias.detach(),
                               m.running_mean if now else None, m.running_var if now else None,
                               BN_MOMENTUM if m.momentum is None else m.momentum, m.eps, True, scratch["mi"], scratch["ss"],
                               out1, act1, out2, act2)
            if use_running and deferred is not None:
                deferred.append((self, scratch["acc"], n * h * w))
            if use_running and m.num_batches_tracked is not None and not self.shared_counter and deferred is None:
                m.num_batches_tracked.add_(1)
        else:
            ops.bn_fused_apply(y, None, n * h * w, m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var,
                               0.0, m.eps, False, scratch["mi"], scratch["ss"], out1, act1, out2, act2)

    def apply_running_update(self, acc, count):
        m = self.m
        ops.bn_running_update(acc, count, m.running_mean, m.running_var, BN_MOMENTUM if m.momentum is None else m.momentum)
        if m.num_batches_tracked is not None and not self.shared_counter:
            m.num_batches_tracked.add_(1)

    def backward(self, y, scratch, training, g1, act1, g2, act2, dy, want_param_grads, zero_acc=True):
        if zero_acc:
            scratch["acc"].zero_()
        ops.bn_act_bwd(y, scratch["ss"], scratch["mi"], self.m.weight.detach(), training, g1, act1, g2, act2,
                       scratch["acc"], dy, self.ggamma if want_param_grads else None,
                       self.gbeta if want_param_grads else None)


def _bn_scratch(c, device):
    return {"acc": torch.zeros((_lib.BN_SLOTS, 2, c), dtype=torch.float64, device=device),
            "mi": torch.empty((2, c), dtype=torch.float32, device=device),
            "ss": torch.empty((2, c), dtype=torch.float32, device=device)}


class _BNArena:
    """Per-pass BatchNorm scratch for a whole network: ONE zero-fill instead of one per layer.  Every layer owns
    BN_SLOTS x [2, C] fp64 accumulators (forward: partial statistic slots; backward: slot 0 holds the two reductions)."""

    def __init__(self, channels, device):
        tot = sum(channels)
        self.acc = torch.zeros(_lib.BN_SLOTS * 2 * tot, dtype=torch.float64, device=device)
        self.f32 = torch.empty(4 * tot, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, c):
        o = self.off
        self.off += c
        k = _lib.BN_SLOTS * 2
        return {"acc": self.acc[k * o:k * (o + c)].view(_lib.BN_SLOTS, 2, c), "mi": self.f32[4 * o:4 * o + 2 * c].view(2, c),
                "ss": self.f32[4 * o + 2 * c:4 * o + 4 * c].view(2, c)}


class _NetRuntimeBase:
    """Holds ConvOps/BNOps of one network, the flat packed-gradient buffer and the Adam views."""

    def __init__(self, convs, bns, precision):
        self.convs, self.bns = convs, bns
        self.precision = precision
        self.act_dtype = ops.DTYPES[precision][1]
        self.flat_grad = None
        self.param_grad_views = {}     # id(param) -> (view, d0, d1)  (d0 = 0 for non-packed gradients)
        # num_batches_tracked of every BatchNorm becomes a view into one int64 tensor: one increment per forward pass
        self.counters = None
        tracked = [b for b in bns if b.m.track_running_stats and b.m.num_batches_tracked is not None]
        if tracked and len(tracked) == len(bns):
            dev = tracked[0].m.num_batches_tracked.device
            self.counters = torch.stack([b.m.num_batches_tracked.detach().reshape(()) for b in tracked]).to(dev)
            for i, b in enumerate(tracked):
                b.m.num_batches_tracked = self.counters[i]
                b.shared_counter = True

    def device(self):
        return self.convs[0].weight.device

    # ---- weight gradients on a side stream ----------------------------------------------------------------------
    # A layer's wgrad only feeds the optimiser, while its dgrad feeds the rest of the backward chain: with `side_stream`
    # set (the engine does), every wgrad launch is forked onto that stream right after the kernels that produced its
    # operands, and the backward pass joins it at its end.  The two kernel families then share the SMs, which fills the
    # partial waves and per-kernel ramps each of them leaves idle on its own.  Works in plain streams and under CUDA-graph
    # capture (fork / join become graph edges).  wgrad kernels only accumulate into the preallocated flat gradient buffer
    # (no allocation on the side stream); their operand tensors are kept alive until the join (`keep`).
    side_stream = None

    def _wgrad_async(self, ws, fn, *tensors):
        side = self.side_stream
        if side is None:
            fn()
            return
        ws.setdefault("keep", []).extend(t for t in tensors if t is not None)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        ws["forked"] = True

    def _join_side(self, ws):
        if ws.pop("forked", False):
            torch.cuda.current_stream().wait_stream(self.side_stream)
        ws.pop("keep", None)

    def pack_sources(self, sources):
        """The torch.cat of cgan.py:281-289,321-324 as ONE packed NHWC tensor: ("b", zero-bordered 8-channel tensor) when
        the first layer runs on the thin tensor-core path, else ("p", cpad-channel tensor).  The result only depends on the
        sources, so the engine packs each distinct concatenation once per step and shares it between networks."""
        self.ensure_packed()
        first = self.convs[0]
        if first.thin == "cin":
            return ("b", ops.pack_input(sources, 8, self.act_dtype, border=1))
        return ("p%d" % self.cpad, ops.pack_input(sources, self.cpad, self.act_dtype))

    def _packed_input(self, sources, packed):
        kind, t = packed if packed is not None else self.pack_sources(sources)
        want = "b" if self.convs[0].thin == "cin" else "p%d" % self.cpad
        if kind != want:
            raise ValueError(f"packed input layout {kind!r} does not fit this network (needs {want!r})")
        return (None, t) if kind == "b" else (t, None)

    def ensure_packed(self, force=False):
        for c in self.convs:
            c.ensure_packed(self.act_dtype, force)

    def alloc_grads(self):
        """One flat fp32 buffer: packed conv-weight gradients, then bias / BN gradients."""
        if self.flat_grad is not None and self.flat_grad.device == self.device():
            return
        sizes = []
        order = getattr(self, "_grad_order", self.convs)     # (a generator puts its up-conv gradients first: bucket "ups")
        for c in order:
            sizes.append(("w", c, 16 * c.d0 * c.d1))
        for c in self.convs:
            if c.bias is not None:
                sizes.append(("b", c, c.cout))
        for b in self.bns:
            sizes.append(("gamma", b, b.c))
            sizes.append(("beta", b, b.c))
        total = sum((s + 3) // 4 * 4 for _, _, s in sizes)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device())
        off = 0
        for kind, obj, s in sizes:
            v = self.flat_grad[off:off + s]
            if kind == "w":
                obj.g = v
                self.param_grad_views[id(obj.weight)] = (v, obj.d0, obj.d1)
            elif kind == "b":
                obj.gb = v
                self.param_grad_views[id(obj.bias)] = (v, 0, 0)
            elif kind == "gamma":
                obj.ggamma = v
                self.param_grad_views[id(obj.m.weight)] = (v, 0, 0)
            else:
                obj.gbeta = v
                self.param_grad_views[id(obj.m.bias)] = (v, 0, 0)
            off += (s + 3) // 4 * 4

    def zero_grads(self):
        self.alloc_grads()
        self.flat_grad.zero_()

    def grad_bucket(self, name=None):
        """Slices of the flat gradient buffer that are all-reduced separately: None / "all" = everything; for a generator
        "ups" = the up-conv weight gradients (final after the decoder half of backward), "rest" = everything else."""
        self.alloc_grads()
        if name in (None, "all"):
            return self.flat_grad
        n_ups, n_deep = getattr(self, "_n_ups", 0), getattr(self, "_n_deep", 0)
        if name == "ups":
            return self.flat_grad[:n_ups]
        if name == "rest":
            return self.flat_grad[n_ups:]
        if name == "deep":
            return self.flat_grad[n_ups:n_ups + n_deep]
        if name == "tail":
            return self.flat_grad[n_ups + n_deep:]
        raise KeyError(name)

    def grads_in_parameter_layout(self, params):
        """torch-layout gradients for autograd (conv weights are un-packed by a kernel)."""
        out = []
        for p in params:
            v, d0, d1 = self.param_grad_views[id(p)]
            out.append(ops.unpack_grad(v, d0, d1) if d0 > 0 else v.clone().view(p.shape))
        return out


# =================================================================================================
# Generator
# =================================================================================================
class GeneratorRuntime(_NetRuntimeBase):
    def __init__(self, downs, down_bns, ups, up_bns, in_channels, out_channels, precision):
        """downs[k-1], ups[k-1]: ConvOps of level k (1 = outermost); down_bns/up_bns: BNOp or None per level."""
        self.L = len(downs)
        self.downs, self.ups, self.down_bns, self.up_bns = downs, ups, down_bns, up_bns
        self.cin, self.cout = in_channels, out_channels
        self.cpad = max(4, (in_channels + 3) // 4 * 4)
        convs = list(downs) + list(ups)
        bns = [b for b in down_bns if b is not None] + [b for b in up_bns if b is not None]
        # flat gradient buffer = [up-conv weights | deep down-conv weights (levels >= DEEP_LEVEL, innermost first: the order in
        # which the encoder half of the backward pass completes them) | the other down convs | biases, BatchNorm]: three
        # buckets ("ups", "deep", "tail") that become final -- and can go on the wire / into Adam -- one after the other
        self.deep_level = min(5, self.L)
        deep = [downs[k - 1] for k in range(self.L, self.deep_level - 1, -1)]
        shallow = [downs[k - 1] for k in range(self.deep_level - 1, 0, -1)]
        self.deep_convs = deep
        self._grad_order = list(ups) + deep + shallow
        pad4 = lambda c: (16 * c.d0 * c.d1 + 3) // 4 * 4
        self._n_ups = sum(pad4(c) for c in ups)
        self._n_deep = sum(pad4(c) for c in deep)
        super().__init__(convs, bns, precision)

    def sizes(self, h, w):
        """spatial size of every level: s[0] = input, s[k] = output of down conv k."""
        s = [(h, w)]
        for k in range(1, self.L + 1):
            ph, pw = s[-1]
            if k > 1:
                ph, pw = ph + ph % 2, pw + pw % 2        # F.pad to even (stcgan_g.py:126-130)
            s.append(self.downs[k - 1].out_size(ph, pw))
        return s

    def forward(self, sources, training, packed=None):
        """sources: list of NCHW fp32 tensors concatenated along channels (the torch.cat of cgan.py:286).
        `packed`: result of pack_sources(sources), if the caller already has it.  Returns (out NCHW fp32, workspace)."""
        self.ensure_packed()
        dt, dev, L = self.act_dtype, self.device(), self.L
        n, _, h, w = sources[0].shape
        # thin first layer: zero-bordered 8-channel input, the thin-K tensor-core path reads its windows from it
        inp, inp_b = self._packed_input(sources, packed)
        thin_in = inp_b is not None
        s = self.sizes(h, w)
        arena = _BNArena([b.c for b in self.bns], dev)
        if training and self.counters is not None:
            self.counters.add_(1)
        ws = {"inp": inp, "inp_b": inp_b, "n": n, "s": s, "training": training, "arena": arena, "y": [None] * (L + 1), "a": [None] * (L + 1),
              "cat": [None] * (L + 1), "uy": [None] * (L + 2), "bn_d": [None] * (L + 1), "bn_u": [None] * (L + 2)}
        C = [None] + [d.cout for d in self.downs]
        new = lambda hh, ww, c: torch.empty((n, hh, ww, c), dtype=dt, device=dev)
        # ---- encoder
        x = None if thin_in else inp[..., :self.cin]
        for k in range(1, L + 1):
            hk, wk = s[k]
            down = self.downs[k - 1]
            sc, fuse = None, False
            if k < L and self.down_bns[k - 1] is not None:
                sc = arena.take(C[k])
                fuse = training and down.stats_fusable()
            y = down.forward(x, hk, wk, x_bordered=inp_b if k == 1 else None, bn_acc=sc["acc"] if fuse else None)
            ws["y"][k] = y
            if k < L:
                a = new(hk, wk, C[k])
                cat = new(hk, wk, 2 * C[k])
                ws["a"][k], ws["cat"][k] = a, cat
                if self.down_bns[k - 1] is not None:
                    ws["bn_d"][k] = sc
                    self.down_bns[k - 1].forward(y, sc, training, a, ACT_LEAKY, cat[..., :C[k]], ACT_RELU, stats_done=fuse)
                else:
                    ops.bn_act_apply(y, None, a, ACT_LEAKY, cat[..., :C[k]], ACT_RELU)
                x = a
            else:
                r = new(hk, wk, C[k])
                ws["a"][k] = r
                ops.bn_act_apply(y, None, r, ACT_RELU)
                x = r
        # ---- decoder
        for k in range(L, 1, -1):
            hk, wk = s[k]
            up = self.ups[k - 1]
            sc = arena.take(up.cout)
            fuse = training and up.stats_fusable()
            uy = up.forward(x, 2 * hk, 2 * wk, bn_acc=sc["acc"] if fuse else None)
            ws["uy"][k] = uy
            ws["bn_u"][k] = sc
            dst = ws["cat"][k - 1]
            self.up_bns[k - 1].forward(uy, sc, training, dst[..., C[k - 1]:], ACT_RELU, stats_done=fuse)   # crops to dst's H, W
            x = dst
        h1, w1 = s[1]
        out = torch.empty((n, self.cout, 2 * h1, 2 * w1), dtype=torch.float32, device=dev)
        self.ups[0].forward(x, 2 * h1, 2 * w1, out_nchw=out, act=ACT_TANH)
        ws["out"] = out
        return out, ws

    def forward_inference(self, sources, packed=None, quantize=False, want_float=True):
        """Eval-mode forward that keeps nothing for a backward pass (CGAN.infer, src/cgan.py:422-438).  `quantize`: the last
        layer's epilogue also writes the uint8 HWC image of utils.float2uint (cgan.py:441-446) and (out, u8) is returned;
        `want_float=False` then skips the float output altogether (out is None).  On the bf16 tensor-core
        path every BatchNorm(eval) + LeakyReLU / ReLU is folded into the producing convolution's epilogue (per-channel scale /
        shift from the running statistics, stcgan_tapconv_ep), the encoder writes both activations of a level -- LeakyReLU for
        the next down conv, ReLU into the left half of the decoder's concat buffer -- from one accumulator, and the decoder's
        odd-size crop (stcgan_g.py:131) is a store bound: no pre-activation tensor is written or re-read, no BatchNorm pass
        runs.  Falls back to forward(training=False) where a layer is not eligible (fp32 mode, narrow networks)."""
        self.ensure_packed()
        L = self.L
        fold_ok = (self.act_dtype == torch.bfloat16 and self.downs[0].thin == "cin" and self.ups[0].thin == "coutT"
                   and all(c.use_tc and c.thin is None and ops.tc_eligible_conv(c.cin, c.cout)
                           for c in list(self.downs[1:]) + list(self.ups[1:])))
        inp, inp_b = self._packed_input(sources, packed)
        if not fold_ok or inp_b is None:
            out = self.forward(sources, False, packed=packed)[0]
            return (out, ops.float2uint_hwc(out)) if quantize else out
        dt, dev = self.act_dtype, self.device()
        n, _, h, w = sources[0].shape
        s = self.sizes(h, w)
        C = [None] + [d.cout for d in self.downs]
        new = lambda hh, ww, c: torch.empty((n, hh, ww, c), dtype=dt, device=dev)

        def folded(bn):             # eval-mode BatchNorm as per-channel (scale, shift): fp32 [2, C]
            m = bn.m
            ss = torch.empty((2, bn.c), dtype=torch.float32, device=dev)
            mi = torch.empty((2, bn.c), dtype=torch.float32, device=dev)
            ops.bn_finalize(None, 0, m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var, 0.0, m.eps, False, mi, ss)
            return ss

        cats = [None] * (L + 1)
        # ---- encoder
        x = None
        for k in range(1, L + 1):
            hk, wk = s[k]
            down = self.downs[k - 1]
            if k == L:              # innermost: no BatchNorm, ReLU feeds the up conv
                x = ops.tapconv_ep(GEOM_WIN_S2, x, down.p1, C[k], hk, wk, act=ACT_RELU)
                break
            a, cat = new(hk, wk, C[k]), new(hk, wk, 2 * C[k])
            cats[k] = cat
            if k == 1:              # outermost: no BatchNorm; thin-K first layer, two activations of one accumulator
                ops.thinconv2(inp_b, 2, down.wthin, C[k], hk, wk, act=ACT_LEAKY, out=a, act2=ACT_RELU, out2=cat[..., :C[k]])
            else:
                ops.tapconv_ep(GEOM_WIN_S2, x, down.p1, C[k], hk, wk, scale_shift=folded(self.down_bns[k - 1]), act=ACT_LEAKY,
                               out=a, act2=ACT_RELU, out2=cat[..., :C[k]])
            x = a
        # ---- decoder
        for k in range(L, 1, -1):
            hk, wk = s[k]
            up = self.ups[k - 1]
            dst = cats[k - 1]
            ops.tapconv_ep(GEOM_PARITY, x, up.p2, up.cout, 2 * hk, 2 * wk, scale_shift=folded(self.up_bns[k - 1]), act=ACT_RELU,
                           out=dst[..., C[k - 1]:], crop=(dst.shape[1], dst.shape[2]))
            x = dst
        h1, w1 = s[1]
        up = self.ups[0]
        out = torch.empty((n, self.cout, 2 * h1, 2 * w1), dtype=torch.float32, device=dev) if (want_float or not quantize) else None
        if quantize:
            u8 = ops.thin_convT_u8(x, up.wtn, up.cpad, up.cout, 2 * h1, 2 * w1, bias=None if up.bias is None else up.bias.detach(),
                                   act=ACT_TANH, out_nchw=out)
            return out, u8
        up.forward(x, 2 * h1, 2 * w1, out_nchw=out, act=ACT_TANH)
        return out

    def backward(self, ws, dout, need_input_grad, param_grads=True, part=None, on_deep=None):
        """dout: NCHW fp32 gradient of the output.  Accumulates parameter gradients into the flat buffer;
        returns the NHWC gradient of the packed input (or None).
        `part`: None = the whole pass; "dec" = the decoder half only (afterwards every up-conv weight gradient -- the
        "ups" gradient bucket, 64 % of the network's parameters -- is final, so its all-reduce can run under the encoder half);
        "enc" = the rest (dout is ignored).  `on_deep` (encoder half): called -- with the stream the weight-gradient kernels
        run on as the current stream -- right after the last weight gradient of the "deep" bucket has been issued."""
        dt, dev, L, s = self.act_dtype, self.device(), self.L, ws["s"]
        training = ws["training"]
        n = ws["n"]
        C = [None] + [d.cout for d in self.downs]
        new = lambda hh, ww, c: torch.empty((n, hh, ww, c), dtype=dt, device=dev)
        if part == "enc":
            dcat = ws.pop("_dcat_last")
            return self._backward_encoder(ws, dcat, need_input_grad, param_grads, on_deep)
        ws["arena"].acc.zero_()            # all BatchNorm backward reductions of this pass accumulate into it
        # ---- outermost up conv (+Tanh)
        up = self.ups[0]
        if up.thin == "coutT":
            g_b = ops.out_act_bwd(ACT_TANH, ws["out"], dout, dt, cpad=8, border=1)
            if param_grads:
                self._wgrad_async(ws, lambda: up.wgrad(ws["cat"][1], None, g_bordered=g_b), g_b)
            dcat = up.dgrad(None, *s[1], g_bordered=g_b)
        else:
            g = ops.out_act_bwd(ACT_TANH, ws["out"], dout, dt)
            if param_grads:
                self._wgrad_async(ws, lambda: up.wgrad(ws["cat"][1], g), g)
            dcat = up.dgrad(g, *s[1])
        # ---- decoder, outside-in
        for k in range(2, L + 1):
            up, bn, sc = self.ups[k - 1], self.up_bns[k - 1], ws["bn_u"][k]
            uy = ws["uy"][k]
            guy = torch.empty_like(uy)
            bn.backward(uy, sc, training, dcat[..., C[k - 1]:], ACT_RELU, None, ACT_NONE, guy, param_grads, zero_acc=False)
            x_in = ws["cat"][k] if k < L else ws["a"][L]
            if param_grads:
                self._wgrad_async(ws, lambda up=up, x_in=x_in, guy=guy: up.wgrad(x_in, guy), guy)
            dcat_prev, dcat = dcat, up.dgrad(guy, *s[k])
            ws.setdefault("dcat", {})[k - 1] = dcat_prev
        if part == "dec":
            ws["_dcat_last"] = dcat
            self._join_side(ws)
            return None
        return self._backward_encoder(ws, dcat, need_input_grad, param_grads, on_deep)

    def _backward_encoder(self, ws, dcat, need_input_grad, param_grads, on_deep=None):
        dt, dev, L, s = self.act_dtype, self.device(), self.L, ws["s"]
        training, n = ws["training"], ws["n"]
        C = [None] + [d.cout for d in self.downs]
        dcats = ws["dcat"]
        # ---- encoder, inside-out.  `da` = gradient w.r.t. the activated tensor feeding the next down conv
        da = dcat                       # at k = L: gradient w.r.t. relu(x_L)
        for k in range(L, 0, -1):
            y = ws["y"][k]
            gy = torch.empty_like(y)
            if k == L:
                ops.bn_act_bwd(y, None, None, None, training, da, ACT_RELU, None, ACT_NONE, None, gy, None, None)
            else:
                skip_g = dcats[k][..., :C[k]]
                bn = self.down_bns[k - 1]
                if bn is not None:
                    bn.backward(y, ws["bn_d"][k], training, da, ACT_LEAKY, skip_g, ACT_RELU, gy, param_grads, zero_acc=False)
                else:
                    ops.bn_act_bwd(y, None, None, None, training, da, ACT_LEAKY, skip_g, ACT_RELU, None, gy, None, None)
            down = self.downs[k - 1]
            thin_in = k == 1 and ws["inp_b"] is not None
            x_in = ws["a"][k - 1] if k > 1 else (None if thin_in else ws["inp"][..., :self.cin])
            if param_grads:
                self._wgrad_async(ws, lambda down=down, x_in=x_in, gy=gy, thin_in=thin_in:
                                  down.wgrad(x_in, gy, x_bordered=ws["inp_b"] if thin_in else None), gy)
            if on_deep is not None and k == self.deep_level:
                import contextlib
                with (torch.cuda.stream(self.side_stream) if (self.side_stream is not None and param_grads)
                      else contextlib.nullcontext()):
                    on_deep()
            if k > 1:
                da = down.dgrad(gy, *s[k - 1])
            elif need_input_grad:
                if thin_in:
                    dinp = torch.empty((n, s[0][0], s[0][1], 8), dtype=dt, device=dev)
                    down.dgrad(gy, *s[0], out8=dinp)
                else:
                    dinp = torch.empty((n, s[0][0], s[0][1], self.cpad), dtype=dt, device=dev)
                    down.dgrad(gy, *s[0], out=dinp[..., :self.cin])
                self._join_side(ws)
                return dinp
        self._join_side(ws)
        return None


# =================================================================================================
# Discriminator
# =================================================================================================
class DiscriminatorRuntime(_NetRuntimeBase):
    def __init__(self, convs, bns, in_channels, use_sigmoid, precision):
        """convs: ConvOps in order; bns[i]: BNOp or None following conv i."""
        self.layers, self.layer_bns = convs, bns
        self.cin = in_channels
        self.cpad = max(4, (in_channels + 3) // 4 * 4)
        self.use_sigmoid = use_sigmoid
        super().__init__(list(convs), [b for b in bns if b is not None], precision)

    def forward(self, sources, training, packed=None, defer_running=False):
        """`defer_running` (training only): BatchNorm running statistics are left untouched; the workspace records the pending
        updates and `apply_deferred_running(ws)` performs them -- see BNOp.forward."""
        self.ensure_packed()
        dt, dev = self.act_dtype, self.device()
        n, _, h, w = sources[0].shape
        inp, inp_b = self._packed_input(sources, packed)
        thin_in = inp_b is not None
        nl = len(self.layers)
        deferred = [] if (defer_running and training) else None
        arena = _BNArena([b.c for b in self.bns], dev)
        if training and self.counters is not None and deferred is None:      # (a deferred pass counts when its updates land)
            self.counters.add_(1)
        ws = {"inp": inp, "inp_b": inp_b, "n": n, "training": training, "arena": arena, "y": [None] * nl, "a": [None] * nl,
              "bn": [None] * nl, "s": [(h, w)], "deferred": deferred}
        x = None if thin_in else inp[..., :self.cin]
        for i, conv in enumerate(self.layers):
            oh, ow = conv.out_size(*ws["s"][-1])
            if oh < 1 or ow < 1:
                raise RuntimeError(f"input {h}x{w} too small for the PatchGAN discriminator")
            ws["s"].append((oh, ow))
            if i == nl - 1:
                out = torch.empty((n, conv.cout, oh, ow), dtype=torch.float32, device=dev)
                conv.forward(x, oh, ow, out_nchw=out, act=ACT_SIGMOID if self.use_sigmoid else ACT_NONE)
                ws["out"] = out
                return out, ws
            bn = self.layer_bns[i]
            if bn is None:
                # bias + LeakyReLU fused into the conv epilogue; sign(a) == sign(pre-activation) serves the backward
                a = conv.forward(x, oh, ow, act=ACT_LEAKY, x_bordered=inp_b if i == 0 else None)
                ws["a"][i] = a
            else:
                sc = arena.take(conv.cout)
                fuse = training and conv.stats_fusable()
                y = conv.forward(x, oh, ow, bn_acc=sc["acc"] if fuse else None)
                a = torch.empty_like(y)
                ws["y"][i], ws["a"][i], ws["bn"][i] = y, a, sc
                bn.forward(y, sc, training, a, ACT_LEAKY, stats_done=fuse, deferred=deferred)
            x = a

    def apply_deferred_running(self, ws):
        """the running-statistics updates a forward(..., defer_running=True) left pending (must run before ws's backward,
        which re-uses the statistic slots)"""
        if ws.get("deferred") is None:
            return
        for bn, acc, count in ws["deferred"]:
            bn.apply_running_update(acc, count)
        if self.counters is not None:
            self.counters.add_(1)
        ws["deferred"] = None

    def backward(self, ws, dout, need_input_grad, param_grads=True):
        dt, dev = self.act_dtype, self.device()
        training, s = ws["training"], ws["s"]
        nl = len(self.layers)
        n = ws["n"]
        last = self.layers[nl - 1]
        ws["arena"].acc.zero_()
        oact = ACT_SIGMOID if self.use_sigmoid else ACT_NONE
        g_b = None
        if last.thin == "cout1":
            g_b = ops.out_act_bwd(oact, ws["out"], dout, dt, cpad=8, border=2)
            g = None
        else:
            g = ops.out_act_bwd(oact, ws["out"], dout, dt)
        for i in range(nl - 1, -1, -1):
            conv = self.layers[i]
            bias_done = False
            thin_in = i == 0 and ws["inp_b"] is not None
            x_in = ws["a"][i - 1] if i > 0 else (None if thin_in else ws["inp"][..., :self.cin])
            if i < nl - 1:
                bn = self.layer_bns[i]
                if bn is None:
                    gy = torch.empty_like(ws["a"][i])
                    fuse_bias = param_grads and thin_in and conv.bias is not None and conv.gb is not None
                    ops.bn_act_bwd(ws["a"][i], None, None, None, training, g, ACT_LEAKY, None, ACT_NONE, None, gy, None, None,
                                   dbias=conv.gb if fuse_bias else None)
                    bias_done = fuse_bias
                else:
                    gy = torch.empty_like(ws["y"][i])
                    bn.backward(ws["y"][i], ws["bn"][i], training, g, ACT_LEAKY, None, ACT_NONE, gy, param_grads, zero_acc=False)
            else:
                gy = g
            if i == nl - 1 and g_b is not None:
                if param_grads:
                    self._wgrad_async(ws, lambda conv=conv, x_in=x_in: conv.wgrad(x_in, None, g_bordered=g_b), g_b)
                g = conv.dgrad(None, *s[i], g_bordered=g_b)
                continue
            if param_grads:
                self._wgrad_async(ws, lambda conv=conv, x_in=x_in, gy=gy, thin_in=thin_in, bias_done=bias_done:
                                  conv.wgrad(x_in, gy, x_bordered=ws["inp_b"] if thin_in else None, bias_done=bias_done), gy)
            if i > 0:
                g = conv.dgrad(gy, *s[i])
            elif need_input_grad:
                if thin_in:
                    dinp = torch.empty((n, s[0][0], s[0][1], 8), dtype=dt, device=dev)
                    conv.dgrad(gy, *s[0], out8=dinp)
                else:
                    dinp = torch.empty((n, s[0][0], s[0][1], self.cpad), dtype=dt, device=dev)
                    conv.dgrad(gy, *s[0], out=dinp[..., :self.cin])
                self._join_side(ws)
                return dinp
        self._join_side(ws)
        return None
