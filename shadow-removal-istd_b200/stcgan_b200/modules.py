"""Drop-in nn.Modules for the reference's ST-CGAN networks.

`UnetGenerator` and `NLayerDiscriminator` keep the reference's constructor signatures
(src/models/stcgan_g.py:12-17, src/models/stcgan_d.py:10-15 -- unknown kwargs are swallowed like the
reference does), forward signatures (NCHW fp32 in, NCHW fp32 out), class names (used in checkpoint file
names, src/cgan.py:475-488), `state_dict()` key layout / `named_parameters()` order (82 and 22 entries), and
the RNG consumption order of the constructors, so `torch.manual_seed(s); UnetGenerator(...)` yields the same
initial weights as the reference.  The torch parameter-holder submodules (nn.Conv2d, nn.ConvTranspose2d,
nn.BatchNorm2d) are never *executed*: forward/backward run the hand-written sm_100a kernels through one
autograd.Function per network.  There is no CPU / eager fallback: non-CUDA input raises.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import nets

_DEFAULT_PRECISION = os.environ.get("STCGAN_B200_PRECISION", "bf16")


def _check_precision(p):
    if p not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {p!r}")
    return p


class _NetFunction(torch.autograd.Function):
    """Whole-network autograd node: forward runs the kernel schedule, backward the hand-derived one."""

    @staticmethod
    def forward(ctx, module, x, *params):
        rt = module.runtime()
        training = module.training
        out, ws = rt.forward([x.detach().contiguous()], training)
        ctx.module, ctx.rt, ctx.ws = module, rt, ws
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        rt, ws = ctx.rt, ctx.ws
        need_x = ctx.needs_input_grad[1]
        need_p = any(ctx.needs_input_grad[2:])
        if need_p:
            rt.zero_grads()
        dinp = rt.backward(ws, dout.contiguous().float(), need_x, param_grads=need_p)
        dx = None
        if need_x:
            n, h, w, _ = dinp.shape
            dx = torch.empty((n, rt.cin, h, w), dtype=torch.float32, device=dout.device)
            from . import ops
            ops.unpack_input_grad(dinp, 0, rt.cin, dx, False)
        pg = [None] * len(ctx.params)
        if need_p:
            grads = rt.grads_in_parameter_layout(ctx.params)
            pg = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[2:])]
        ctx.ws = None
        return (None, dx, *pg)


class _KernelBackedNet(nn.Module):
    """Shared plumbing: precision selection, runtime cache, loud failure off-GPU."""

    def _init_backend(self, precision):
        self._precision = _check_precision(precision or _DEFAULT_PRECISION)
        self._rt = None

    @property
    def precision(self):
        return self._precision

    def set_precision(self, precision):
        self._precision = _check_precision(precision)
        self._rt = None
        return self

    def runtime(self):
        dev = next(self.parameters()).device
        if self._rt is None or self._rt.device() != dev or self._rt.precision != self._precision:
            self._rt = self._build_runtime()
        return self._rt

    def _apply(self, fn, *a, **kw):       # .to()/.cuda()/.double() invalidate packed copies
        self._rt = None
        return super()._apply(fn, *a, **kw)

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError(f"{type(self).__name__} (stcgan_b200) runs on CUDA only: hand-written sm_100a kernels, "
                               "no CPU fallback")
        if input.dtype != torch.float32:
            raise TypeError("expected a float32 NCHW tensor, as produced by the reference's ISTDDataset")
        p = next(self.parameters())
        if p.device != input.device or p.dtype != torch.float32:
            raise RuntimeError("module parameters must be float32 on the input's device (call .to(device))")
        return _NetFunction.apply(self, input, *self.parameters())


class UnetLevel(nn.Module):
    """Parameter container mirroring one reference skip block: only `self.model` (a Sequential whose index
    layout matches src/models/stcgan_g.py:96-118) matters."""

    def __init__(self, items):
        super().__init__()
        self.model = nn.Sequential(*items)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("UnetLevel is a parameter container; call the UnetGenerator")


class UnetGenerator(_KernelBackedNet):
    """8-level pix2pix U-Net (reference: src/models/stcgan_g.py:9-57)."""

    def __init__(self, in_channels, out_channels, ngf=64, num_downs=8, norm_layer=nn.BatchNorm2d,
                 use_dropout=False, precision=None, **kwargs):
        super().__init__()
        if norm_layer is not nn.BatchNorm2d:
            raise NotImplementedError("stcgan_b200 implements the BatchNorm2d configuration the reference uses")
        if use_dropout:
            raise NotImplementedError("use_dropout is unreachable from src/cgan.py and not implemented")
        if num_downs < 5:
            raise ValueError("num_downs must be >= 5 (reference topology)")
        self.in_channels, self.out_channels, self.ngf, self.num_downs = in_channels, out_channels, ngf, num_downs
        widths = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)
        # build innermost -> outermost so the global RNG is consumed in the reference's order
        downs, ups, dnorm, unorm = {}, {}, {}, {}
        block = None
        for k in range(num_downs, 0, -1):
            d_in = in_channels if k == 1 else widths[k - 2]
            d_out = widths[k - 1]
            u_in = d_out if k == num_downs else 2 * d_out
            u_out = out_channels if k == 1 else widths[k - 2]
            down = nn.Conv2d(d_in, d_out, kernel_size=4, stride=2, padding=1, bias=False)
            bn_d = nn.BatchNorm2d(d_out)          # constructed (like the reference) even where unused
            bn_u = nn.BatchNorm2d(u_out)
            up = nn.ConvTranspose2d(u_in, u_out, kernel_size=4, stride=2, padding=1, bias=(k == 1))
            downs[k], ups[k] = down, up
            if k == 1:
                items = [down, block, nn.ReLU(True), up, nn.Tanh()]
                dnorm[k], unorm[k] = None, None
            elif k == num_downs:
                items = [nn.LeakyReLU(0.2, True), down, nn.ReLU(True), up, bn_u]
                dnorm[k], unorm[k] = None, bn_u
            else:
                items = [nn.LeakyReLU(0.2, True), down, bn_d, block, nn.ReLU(True), up, bn_u]
                dnorm[k], unorm[k] = bn_d, bn_u
            block = UnetLevel(items)
        self.model = block
        self._levels = (downs, dnorm, ups, unorm)     # plain tuple of dicts: not registered twice
        self._init_backend(precision)

    def _build_runtime(self):
        downs, dnorm, ups, unorm = self._levels
        L = self.num_downs
        mk = lambda d: None if d is None else nets.BNOp(d)
        return nets.GeneratorRuntime(
            [nets.ConvOp("conv2", downs[k].weight, None) for k in range(1, L + 1)],
            [mk(dnorm[k]) for k in range(1, L + 1)],
            [nets.ConvOp("convT", ups[k].weight, ups[k].bias) for k in range(1, L + 1)],
            [mk(unorm[k]) for k in range(1, L + 1)],
            self.in_channels, self.out_channels, self._precision)


class NLayerDiscriminator(_KernelBackedNet):
    """70x70 PatchGAN (reference: src/models/stcgan_d.py:9-58)."""

    def __init__(self, in_channels, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False,
                 precision=None, **kwargs):
        super().__init__()
        if norm_layer is not nn.BatchNorm2d:
            raise NotImplementedError("stcgan_b200 implements the BatchNorm2d configuration the reference uses")
        self.in_channels, self.use_sigmoid = in_channels, bool(use_sigmoid)
        seq = [nn.Conv2d(in_channels, ndf, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        self._convs, self._bns = [seq[0]], [None]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            conv = nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=2 if n < n_layers else 1, padding=1, bias=False)
            bn = nn.BatchNorm2d(ndf * mult)
            seq += [conv, bn, nn.LeakyReLU(0.2, True)]
            self._convs.append(conv)
            self._bns.append(bn)
        last = nn.Conv2d(ndf * mult, 1, kernel_size=4, stride=1, padding=1)
        seq.append(last)
        self._convs.append(last)
        self._bns.append(None)
        if use_sigmoid:
            seq.append(nn.Sigmoid())
        self.model = nn.Sequential(*seq)
        self._convs, self._bns = tuple(self._convs), tuple(self._bns)
        self._init_backend(precision)

    def _build_runtime(self):
        convs = [nets.ConvOp("conv2" if c.stride[0] == 2 else "conv1", c.weight, c.bias) for c in self._convs]
        bns = [None if b is None else nets.BNOp(b) for b in self._bns]
        return nets.DiscriminatorRuntime(convs, bns, self.in_channels, self.use_sigmoid, self._precision)
