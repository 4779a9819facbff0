"""Loss modules with the reference's API (src/loss.py:14-26, 59-112), backed by the fused loss kernel.

`cal_loss` keeps the reference's inverted flag (src/loss.py:79-84): ls=False -> MSE against the label,
ls=True -> BCE-with-logits against the label; labels real 1.0, fake 0.0 (ls=False) / -1.0 (ls=True).
Value and gradient come out of ONE kernel launch (`stcgan_fused_loss`); the backward of the autograd
node only scales the stored gradient by the incoming scalar.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _FusedTermFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, a, b, target, scale):
        if not a.is_cuda:
            raise RuntimeError("stcgan_b200 losses run on CUDA only (no CPU fallback)")
        a32 = a.detach().contiguous().float()
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
        grad = torch.empty_like(a32) if ctx.needs_input_grad[1] else None
        term = dict(kind=kind, a=a32, grad=grad, target=target, weight=scale, slot=0)
        if b is not None:
            term["b"] = b.detach().contiguous().float()
        ops.fused_loss([term], out)
        ctx.grad = grad
        ctx.shape = a.shape
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        g = None if ctx.grad is None else (ctx.grad * gout).view(ctx.shape)
        return None, g, None, None, None


class _RelLogitsFn(torch.autograd.Function):
    """a - b (RpGAN) or a - b.mean(dim=0) (RaGAN) as one kernel, with its gradient (src/loss.py:88-96, 102-110)."""

    @staticmethod
    def forward(ctx, a, b, avg):
        if not a.is_cuda:
            raise RuntimeError("stcgan_b200 losses run on CUDA only (no CPU fallback)")
        ctx.avg = avg
        return ops.rel_logits(a.detach().contiguous().float(), b.detach().contiguous().float(), avg)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        ga = g if ctx.needs_input_grad[0] else None
        gb = ops.rel_logits(None, g, ctx.avg, backward=True) if ctx.needs_input_grad[1] else None
        return ga, gb, None


class DataLoss(nn.Module):
    """L1 between prediction and target (src/loss.py:14-26)."""
    __slots__ = ["reduction", "norm"]

    def __init__(self, norm=F.l1_loss, reduction: str = "mean"):
        super().__init__()
        if norm is not F.l1_loss:
            raise NotImplementedError("stcgan_b200.DataLoss implements the L1 norm the reference uses")
        if reduction not in ("mean", "sum"):
            raise NotImplementedError("reduction must be 'mean' or 'sum'")
        self.reduction, self.norm = reduction, norm

    def forward(self, y_pred, y_target):
        if y_pred.shape != y_target.shape:
            raise ValueError("DataLoss: shape mismatch")
        scale = 1.0 if self.reduction == "mean" else float(y_pred.numel())
        return _FusedTermFn.apply(ops.KIND_L1, y_pred, y_target, 0.0, scale)


class AdversarialLoss(nn.Module):
    """SGAN / RpGAN / RaGAN discriminator and generator objectives (src/loss.py:59-112)."""

    def __init__(self, ls=False, rel=False, avg=False):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(1.0))
        self.register_buffer("fake_label", torch.tensor(-1.0 if ls else 0.0))
        self.ls, self.rel, self.avg = ls, rel, avg
        self._labels = (1.0, -1.0 if ls else 0.0)     # host copies: no device sync in the hot loop

    def cal_loss(self, C_out, label):
        target = label if isinstance(label, float) else float(label)
        return _FusedTermFn.apply(ops.KIND_BCE if self.ls else ops.KIND_MSE, C_out, None, target, 1.0)

    def forward(self, C_real, C_fake, D_loss=True):
        real, fake = self._labels
        if D_loss:
            if self.rel:
                if self.avg:   # RaGAN (loss.py:90-94): logits relative to the batch mean of the other side
                    return (self.cal_loss(_RelLogitsFn.apply(C_real, C_fake, True), real)
                            + self.cal_loss(_RelLogitsFn.apply(C_fake, C_real, True), fake)) * 0.5
                return self.cal_loss(_RelLogitsFn.apply(C_real, C_fake, False), real)    # RpGAN (loss.py:96)
            return (self.cal_loss(C_real, real) + self.cal_loss(C_fake, fake)) * 0.5   # SGAN (loss.py:98-100)
        if self.rel:
            if self.avg:
                return (self.cal_loss(_RelLogitsFn.apply(C_fake, C_real, True), real)
                        + self.cal_loss(_RelLogitsFn.apply(C_real, C_fake, True), fake)) * 0.5
            return self.cal_loss(_RelLogitsFn.apply(C_fake, C_real, False), real)
        return self.cal_loss(C_fake, real)                                      # SGAN (loss.py:112)
