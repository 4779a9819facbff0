"""Tensor-level wrappers over the C ABI.  torch is used for device memory and streams only.

Activations are NHWC tensors (`torch.bfloat16` in "bf16" mode, `torch.float32` in "fp32" mode);
a channel slice of a wider NHWC buffer is a normal torch view (`buf[..., :C]`), its pixel pitch
is read from `stride(2)`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, BACKEND_FFMA, BACKEND_TC, BF16, F32,
                   GEOM_PARITY, GEOM_WIN_S1, GEOM_WIN_S1_FLIP, GEOM_WIN_S2, LossTerm, check)

import os as _os
_TRACE = bool(_os.environ.get("STCGAN_TRACE"))

DTYPES = {"fp32": (F32, torch.float32), "bf16": (BF16, torch.bfloat16)}


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stcgan_b200 kernels need CUDA tensors (no CPU fallback exists)")


def _nhwc(t: torch.Tensor):
    """(N, H, W, C, pitch) of an NHWC tensor or channel-slice view."""
    assert t.dim() == 4, "expected an NHWC tensor"
    n, h, w, c = t.shape
    assert c == 1 or t.stride(3) == 1, "expected an NHWC tensor with unit channel stride"
    # strides of size-1 dimensions are arbitrary in torch: derive the pixel pitch from the first dimension that has one
    if w > 1:
        ld = t.stride(2)
    elif h > 1:
        ld = t.stride(1)
    elif n > 1:
        ld = t.stride(0)
    else:
        ld = c
    assert ld >= c, "NHWC view: pixel pitch smaller than the channel count"
    assert (h == 1 or t.stride(1) == w * ld) and (n == 1 or t.stride(0) == h * w * ld), "NHWC view must be pixel-contiguous"
    return n, h, w, c, ld


SPLITK_MAX_ELEMS = 4 * 1024 * 1024     # only small outputs (bottleneck layers) get a split-K workspace
# (elements, device, raw stream handle) -> zeroed fp32 workspace (kept all-zero between calls by the finisher kernel).
# Keyed by the ACTUAL stream of the call: two streams -- engine lanes, a user's inference stream next to a training step, a
# second engine -- can never `red.add` into / clear the same buffer under each other.  The buffers are small (<= 16 MB, a few
# sizes per stream) and live for the life of the process.
_SPLITK_WS = {}


def register_concurrent_stream(stream):
    """Kept for callers of the round-1 API: split-K scratch is now always per stream, nothing to register."""
    return None


def tc_eligible_conv(k: int, nout: int) -> bool:
    return k % 64 == 0 and nout % 64 == 0


def tc_eligible_wgrad(d0: int, d1: int) -> bool:
    return d0 % 128 == 0 and d1 % 64 == 0


def tapconv(geom, x, wp, nout, oh, ow, *, bias=None, act=ACT_NONE, out=None, out_nchw=None, backend=BACKEND_FFMA,
            bn_acc=None):
    """out[N,OH,OW,nout] = tap-GEMM(x, wp).  `out` may be a channel-slice view; `out_nchw` (fp32 [N,nout,OH,OW])
    selects the NCHW epilogue instead.  `bn_acc` (fp64 [BN_SLOTS, 2, nout], zeroed by the caller; tensor-core backend
    only): the following BatchNorm's batch statistics are accumulated by the GEMM epilogue."""
    _need_cuda(x, wp)
    n, ih, iw, k, ldx = _nhwc(x)
    lib = _lib.load()
    if out_nchw is not None:
        y, ldy, nchw = out_nchw, nout, 1
        assert out_nchw.dtype == torch.float32 and out_nchw.is_contiguous()
    else:
        if out is None:
            out = torch.empty((n, oh, ow, nout), dtype=x.dtype, device=x.device)
        _, _, _, _, ldy = _nhwc(out)
        y, nchw = out, 0
    ws, ws_bytes = None, 0
    if _TRACE:
        import sys
        print(f"[stcgan trace] tapconv geom={geom} x={tuple(x.shape)} ldx={ldx} -> ({oh},{ow},{nout}) ldy={ldy} backend={backend} "
              f"bias={bias is not None} act={act} bn_acc={bn_acc is not None} x_ptr={x.data_ptr():#x} y_ptr={y.data_ptr():#x}",
              file=sys.stderr, flush=True)
    if backend == BACKEND_TC and n * oh * ow * nout <= SPLITK_MAX_ELEMS and k >= 256:
        # split-K partial sums: one persistent, self-cleaning (all-zero between calls) workspace per size and stream --
        # no allocation and no memset launch per convolution; concurrent streams never share one
        key = (n * oh * ow * nout, x.device.index, torch.cuda.current_stream().cuda_stream)
        wst = _SPLITK_WS.get(key)
        if wst is None:
            wst = _SPLITK_WS[key] = torch.zeros(key[0], dtype=torch.float32, device=x.device)
        ws, ws_bytes = wst.data_ptr(), -wst.numel() * 4
    if bn_acc is not None:
        assert backend == BACKEND_TC and bias is None and act == ACT_NONE and not nchw
        assert bn_acc.dtype == torch.float64 and bn_acc.is_contiguous() and bn_acc.numel() == _lib.BN_SLOTS * 2 * nout
        check(lib.stcgan_tapconv_bnstats(geom, x.data_ptr(), n, ih, iw, k, ldx, wp.data_ptr(), y.data_ptr(), oh, ow, nout,
                                         ldy, ws, ws_bytes, bn_acc.data_ptr(), _stream()), "stcgan_tapconv_bnstats")
        return y
    check(lib.stcgan_tapconv(geom, _code(x), backend, x.data_ptr(), n, ih, iw, k, ldx, wp.data_ptr(),
                             None if bias is None else bias.data_ptr(), act, y.data_ptr(), oh, ow, nout, ldy, nchw,
                             ws, ws_bytes, _stream()), "stcgan_tapconv")
    return y


def _splitk_ws(n_elems, device):
    """persistent, self-cleaning split-K workspace of this size for the current stream (see _SPLITK_WS)"""
    key = (n_elems, device.index, torch.cuda.current_stream().cuda_stream)
    wst = _SPLITK_WS.get(key)
    if wst is None:
        wst = _SPLITK_WS[key] = torch.zeros(n_elems, dtype=torch.float32, device=device)
    return wst


def tapconv_ep(geom, x, wp, nout, oh, ow, *, scale_shift=None, shift=None, act=ACT_NONE, out=None, act2=ACT_NONE, out2=None,
               crop=None):
    """Inference convolution (bf16 tensor cores): v = acc * scale + shift, out = act(v) [, out2 = act2(v)], stores cropped
    to `crop` = (HC, WC).  `scale_shift`: fp32 [2, nout] as written by bn_finalize (eval-mode BatchNorm folded into the
    epilogue); `shift` alone = a conv bias.  `out` / `out2` may be channel-slice views of [N, HC, WC, *] buffers."""
    _need_cuda(x, wp)
    n, ih, iw, k, ldx = _nhwc(x)
    hc, wc = crop if crop is not None else (oh, ow)
    if out is None:
        out = torch.empty((n, hc, wc, nout), dtype=x.dtype, device=x.device)
    assert tuple(out.shape) == (n, hc, wc, nout) and x.dtype == torch.bfloat16
    _, _, _, _, ldy = _nhwc(out)
    ld2 = 0
    if out2 is not None:
        assert tuple(out2.shape) == (n, hc, wc, nout)
        _, _, _, _, ld2 = _nhwc(out2)
    sc = sh = None
    if scale_shift is not None:
        assert scale_shift.dtype == torch.float32 and scale_shift.is_contiguous() and scale_shift.numel() == 2 * nout
        sc, sh = scale_shift.data_ptr(), scale_shift.data_ptr() + 4 * nout
    elif shift is not None:
        sh = shift.data_ptr()
    ws, ws_bytes = None, 0
    if n * oh * ow * nout <= SPLITK_MAX_ELEMS and k >= 256:
        wst = _splitk_ws(n * oh * ow * nout, x.device)
        ws, ws_bytes = wst.data_ptr(), -wst.numel() * 4
    check(_lib.load().stcgan_tapconv_ep(geom, x.data_ptr(), n, ih, iw, k, ldx, wp.data_ptr(), sc, sh, act, out.data_ptr(), ldy,
                                        act2, None if out2 is None else out2.data_ptr(), ld2, oh, ow, hc, wc, nout, ws, ws_bytes,
                                        _stream()), "stcgan_tapconv_ep")
    return out


def thinconv2(t, stride, wthin, nout, oh, ow, *, bias=None, act=ACT_NONE, out=None, act2=ACT_NONE, out2=None):
    """thin-K convolution with two activated outputs (see stcgan_thinconv2)."""
    n, hp, wp_, c, _ = _nhwc(t)
    assert c == 8 and t.is_contiguous() and t.dtype == torch.bfloat16 and out is not None and out2 is not None
    _, _, _, _, ldy = _nhwc(out)
    _, _, _, _, ld2 = _nhwc(out2)
    check(_lib.load().stcgan_thinconv2(t.data_ptr(), n, hp, wp_, stride, wthin.data_ptr(),
                                       None if bias is None else bias.data_ptr(), act, out.data_ptr(), ldy, act2,
                                       out2.data_ptr(), ld2, oh, ow, nout, _stream()), "stcgan_thinconv2")
    return out


def tapwgrad(geom, s, l, g, *, backend=BACKEND_FFMA):
    """g[16, D0, D1] (fp32) += wgrad(S small-grid tensor, L large-grid tensor)."""
    _need_cuda(s, l, g)
    n, sh, sw, d0, lds = _nhwc(s)
    n2, lh, lw, d1, ldl = _nhwc(l)
    assert n == n2 and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == 16 * d0 * d1
    check(_lib.load().stcgan_tapwgrad(geom, _code(s), backend, s.data_ptr(), n, sh, sw, d0, lds, l.data_ptr(), lh, lw,
                                      d1, ldl, g.data_ptr(), _stream()), "stcgan_tapwgrad")
    return g


def pack_weight(w, p1, p2):
    """w [d0, d1, 4, 4] fp32 -> p1 [16, d0, d1], p2 [16, d1, d0] (either may be None)."""
    _need_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    d0, d1 = w.shape[0], w.shape[1]
    ref = p1 if p1 is not None else p2
    check(_lib.load().stcgan_pack_weight(_code(ref), w.data_ptr(), d0, d1, None if p1 is None else p1.data_ptr(),
                                         None if p2 is None else p2.data_ptr(), _stream()), "stcgan_pack_weight")


def unpack_grad(g, d0, d1, grad=None, accumulate=False):
    _need_cuda(g)
    if grad is None:
        grad = torch.empty((d0, d1, 4, 4), dtype=torch.float32, device=g.device)
    check(_lib.load().stcgan_unpack_grad(g.data_ptr(), d0, d1, grad.data_ptr(), int(accumulate), _stream()),
          "stcgan_unpack_grad")
    return grad


def bn_stats(y, acc):
    n, h, w, c, ld = _nhwc(y)
    check(_lib.load().stcgan_bn_stats(_code(y), y.data_ptr(), n * h * w, c, ld, acc.data_ptr(), _stream()), "stcgan_bn_stats")


def bn_finalize(acc, count, gamma, beta, rmean, rvar, momentum, eps, training, mean_invstd, scale_shift):
    c = gamma.numel()
    check(_lib.load().stcgan_bn_finalize(None if acc is None else acc.data_ptr(), count, c, gamma.data_ptr(), beta.data_ptr(),
                                         None if rmean is None else rmean.data_ptr(),
                                         None if rvar is None else rvar.data_ptr(), momentum, eps, int(training),
                                         mean_invstd.data_ptr(), scale_shift.data_ptr(), _stream()), "stcgan_bn_finalize")


def bn_running_update(acc, count, rmean, rvar, momentum):
    """the momentum update of the running statistics from the fp64 statistic slots `acc` (see stcgan_bn_running_update)."""
    c = rmean.numel()
    assert acc.dtype == torch.float64 and acc.is_contiguous() and acc.numel() == _lib.BN_SLOTS * 2 * c
    check(_lib.load().stcgan_bn_running_update(acc.data_ptr(), int(count), c, rmean.data_ptr(), rvar.data_ptr(), float(momentum),
                                               _stream()), "stcgan_bn_running_update")


def bn_act_apply(y, scale_shift, out1, act1, out2=None, act2=ACT_NONE):
    n, h, w, c, ldy = _nhwc(y)
    n1, hc, wc, c1, ld1 = _nhwc(out1)
    assert c1 == c and n1 == n
    ld2 = 0
    if out2 is not None:
        _, h2, w2, c2, ld2 = _nhwc(out2)
        assert (h2, w2, c2) == (hc, wc, c)
    check(_lib.load().stcgan_bn_act_apply(_code(y), y.data_ptr(), n, h, w, c, ldy,
                                          None if scale_shift is None else scale_shift.data_ptr(), hc, wc,
                                          out1.data_ptr(), ld1, act1, None if out2 is None else out2.data_ptr(), ld2, act2,
                                          _stream()), "stcgan_bn_act_apply")


def bn_fused_apply(y, acc, count, gamma, beta, rmean, rvar, momentum, eps, training, mean_invstd, scale_shift,
                   out1, act1, out2=None, act2=ACT_NONE):
    """bn_finalize + bn_act_apply in one launch; `acc` is fp64 [BN_SLOTS, 2, C] (training) or None (eval)."""
    n, h, w, c, ldy = _nhwc(y)
    n1, hc, wc, c1, ld1 = _nhwc(out1)
    assert c1 == c and n1 == n
    ld2 = 0
    if out2 is not None:
        _, h2, w2, c2, ld2 = _nhwc(out2)
        assert (h2, w2, c2) == (hc, wc, c)
    p = lambda t: None if t is None else t.data_ptr()
    check(_lib.load().stcgan_bn_fused_apply(_code(y), y.data_ptr(), n, h, w, c, ldy, p(acc), count, gamma.data_ptr(),
                                            beta.data_ptr(), p(rmean), p(rvar), momentum, eps, int(training),
                                            mean_invstd.data_ptr(), scale_shift.data_ptr(), hc, wc, out1.data_ptr(), ld1,
                                            act1, p(out2), ld2, act2, _stream()), "stcgan_bn_fused_apply")


def bn_act_bwd(y, scale_shift, mean_invstd, gamma, training, g1, act1, g2, act2, acc, dy, dgamma, dbeta, dbias=None):
    """Two-pass BN(+activation) backward; with scale_shift None this is the plain activation backward.  `acc`: fp64
    [BN_SLOTS, 2, C], zeroed.  `dbias` (fp32 [C]): also accumulates the column sums of dy (bias gradient)."""
    n, h, w, c, ldy = _nhwc(y)
    _, hc, wc, _, ldg1 = _nhwc(g1)
    ldg2 = 0
    if g2 is not None:
        _, h2, w2, _, ldg2 = _nhwc(g2)
        assert (h2, w2) == (hc, wc)
    _, _, _, _, lddy = _nhwc(dy)
    lib = _lib.load()
    p = lambda t: None if t is None else t.data_ptr()
    if (scale_shift is not None and training and dbias is None and y.dtype == torch.bfloat16
            and n * h * w * c <= SMALL_BN_ELEMS and _SMALL_BN):
        # bottleneck-sized tensor: reduce + apply in ONE single-block launch (stcgan_bn_act_bwd_small)
        check(lib.stcgan_bn_act_bwd_small(_code(y), y.data_ptr(), n, h, w, c, ldy, scale_shift.data_ptr(), mean_invstd.data_ptr(),
                                          gamma.data_ptr(), hc, wc, g1.data_ptr(), ldg1, act1, p(g2), ldg2, act2, dy.data_ptr(),
                                          lddy, p(dgamma), p(dbeta), _stream()), "stcgan_bn_act_bwd_small")
        return
    if scale_shift is not None and training:
        check(lib.stcgan_bn_act_bwd_reduce(_code(y), y.data_ptr(), n, h, w, c, ldy, scale_shift.data_ptr(),
                                           mean_invstd.data_ptr(), hc, wc, g1.data_ptr(), ldg1, act1, p(g2), ldg2, act2,
                                           acc.data_ptr(), _stream()), "stcgan_bn_act_bwd_reduce")
    elif scale_shift is not None and dgamma is not None:
        # eval-mode BN: parameter gradients still need the reductions
        check(lib.stcgan_bn_act_bwd_reduce(_code(y), y.data_ptr(), n, h, w, c, ldy, scale_shift.data_ptr(),
                                           mean_invstd.data_ptr(), hc, wc, g1.data_ptr(), ldg1, act1, p(g2), ldg2, act2,
                                           acc.data_ptr(), _stream()), "stcgan_bn_act_bwd_reduce")
    check(lib.stcgan_bn_act_bwd_apply(_code(y), y.data_ptr(), n, h, w, c, ldy, p(scale_shift), p(mean_invstd), p(gamma),
                                      int(training), hc, wc, g1.data_ptr(), ldg1, act1, p(g2), ldg2, act2, p(acc),
                                      dy.data_ptr(), lddy, p(dgamma), p(dbeta), p(dbias), _stream()), "stcgan_bn_act_bwd_apply")


def colsum(g, out):
    n, h, w, c, ld = _nhwc(g)
    check(_lib.load().stcgan_colsum(_code(g), g.data_ptr(), n * h * w, c, ld, out.data_ptr(), _stream()), "stcgan_colsum")


def pack_input(sources, cpad, dtype, border=0):
    """NCHW fp32 sources (<= 3) -> one NHWC tensor with `cpad` channels (zero padded) and an optional zero frame of
    `border` pixels ([N, H+2b, W+2b, cpad])."""
    srcs = [s for s in sources]
    _need_cuda(*srcs)
    n, _, h, w = srcs[0].shape
    for s in srcs:
        assert s.dtype == torch.float32 and s.is_contiguous() and s.shape[0] == n and s.shape[2:] == (h, w)
    out = torch.empty((n, h + 2 * border, w + 2 * border, cpad), dtype=dtype, device=srcs[0].device)
    args = []
    for i in range(3):
        if i < len(srcs):
            args += [srcs[i].data_ptr(), srcs[i].shape[1]]
        else:
            args += [None, 0]
    check(_lib.load().stcgan_pack_input(_code(out), *args, n, h, w, border, out.data_ptr(), cpad, _stream()), "stcgan_pack_input")
    return out


def tapconv_thin_n(geom, x, wp16, nout, oh, ow, *, bias=None, act=ACT_NONE, out8=None, out_nchw=None):
    """thin-N tap-GEMM on the tensor cores (nout <= 16): NCHW fp32 output (bias + activation) or an 8-channel NHWC one."""
    n, ih, iw, k, ldx = _nhwc(x)
    ldy = 0
    if out8 is not None:
        _, _, _, _, ldy = _nhwc(out8)
    check(_lib.load().stcgan_tapconv_thin_n(geom, x.data_ptr(), n, ih, iw, k, ldx, wp16.data_ptr(),
                                            None if bias is None else bias.data_ptr(), act,
                                            None if out8 is None else out8.data_ptr(), ldy,
                                            None if out_nchw is None else out_nchw.data_ptr(), oh, ow, nout, _stream()),
          "stcgan_tapconv_thin_n")


def thin_col2im(mode, x, wt, cpad, cout, oh, ow, *, bias=None, act=ACT_NONE, out_nchw=None, out8=None):
    """thin-N convolution as one pixel GEMM + in-CTA col2im (x is read once): mode 0 = ConvTranspose2d forward / Conv2d-s2
    input gradient, mode 1 = Conv2d(k4,s1,p1) forward.  Output NCHW fp32 (bias + activation) or 8-channel NHWC bf16."""
    _need_cuda(x, wt)
    n, ih, iw, k, ldx = _nhwc(x)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and wt.numel() == 16 * cpad * k
    ldy = 0
    if out8 is not None:
        _, _, _, _, ldy = _nhwc(out8)
    else:
        assert out_nchw.dtype == torch.float32 and out_nchw.is_contiguous() and tuple(out_nchw.shape) == (n, cout, oh, ow)
    check(_lib.load().stcgan_thin_col2im(mode, x.data_ptr(), n, ih, iw, k, ldx, wt.data_ptr(), cpad, cout,
                                         None if bias is None else bias.data_ptr(), act,
                                         None if out_nchw is None else out_nchw.data_ptr(),
                                         None if out8 is None else out8.data_ptr(), ldy, oh, ow, _stream()),
          "stcgan_thin_col2im")


def thin_convT_u8(x, wt, cpad, cout, oh, ow, *, bias=None, act=ACT_TANH, out_nchw=None):
    """the generators' last layer with the uint8 image quantisation in its epilogue (see stcgan_thin_convt_u8): returns the
    uint8 [N, OH, OW, cout] image; `out_nchw` (fp32 [N, cout, OH, OW]) additionally receives the float output."""
    _need_cuda(x, wt)
    n, ih, iw, k, ldx = _nhwc(x)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and wt.numel() == 16 * cpad * k
    if out_nchw is not None:
        assert out_nchw.dtype == torch.float32 and out_nchw.is_contiguous() and tuple(out_nchw.shape) == (n, cout, oh, ow)
    u8 = torch.empty((n, oh, ow, cout), dtype=torch.uint8, device=x.device)
    check(_lib.load().stcgan_thin_convt_u8(x.data_ptr(), n, ih, iw, k, ldx, wt.data_ptr(), cpad, cout,
                                           None if bias is None else bias.data_ptr(), act,
                                           None if out_nchw is None else out_nchw.data_ptr(), u8.data_ptr(), oh, ow, _stream()),
          "stcgan_thin_convt_u8")
    return u8


def pack_weight_tapn(w, n_is_d0, cpad, out):
    d0, d1 = w.shape[0], w.shape[1]
    check(_lib.load().stcgan_pack_weight_tapn(w.data_ptr(), d0, d1, int(n_is_d0), cpad, out.data_ptr(), _stream()),
          "stcgan_pack_weight_tapn")


def thinconv(t, stride, wthin, nout, oh, ow, *, bias=None, act=ACT_NONE, out=None):
    """thin-K convolution of a zero-bordered 8-channel tensor t [N, HP, WP, 8] (tensor cores)."""
    n, hp, wp_, c, _ = _nhwc(t)
    assert c == 8 and t.is_contiguous() and t.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((n, oh, ow, nout), dtype=t.dtype, device=t.device)
    _, _, _, _, ldy = _nhwc(out)
    check(_lib.load().stcgan_thinconv(t.data_ptr(), n, hp, wp_, stride, wthin.data_ptr(),
                                      None if bias is None else bias.data_ptr(), act, out.data_ptr(), oh, ow, nout, ldy,
                                      _stream()), "stcgan_thinconv")
    return out


def thinwgrad(t, stride, thin_c, f, g, fat_is_dim0, flip):
    """g (packed [16][d0][d1] fp32) += thin weight gradient from the zero-bordered thin tensor t and the fat tensor f."""
    n, hp, wp_, c, _ = _nhwc(t)
    n2, fh, fw, dfat, ldf = _nhwc(f)
    assert c == 8 and n == n2 and t.is_contiguous() and g.numel() == 16 * dfat * thin_c
    check(_lib.load().stcgan_thinwgrad(t.data_ptr(), n, hp, wp_, stride, thin_c, f.data_ptr(), fh, fw, dfat, ldf,
                                       int(fat_is_dim0), int(flip), g.data_ptr(), _stream()), "stcgan_thinwgrad")


def pack_weight_thin(w, n_is_d0, flip, out):
    d0, d1 = w.shape[0], w.shape[1]
    check(_lib.load().stcgan_pack_weight_thin(w.data_ptr(), d0, d1, int(n_is_d0), int(flip), out.data_ptr(), _stream()),
          "stcgan_pack_weight_thin")


def pack_weight_pad16(w, n_is_d0, out):
    d0, d1 = w.shape[0], w.shape[1]
    check(_lib.load().stcgan_pack_weight_pad16(w.data_ptr(), d0, d1, int(n_is_d0), out.data_ptr(), _stream()),
          "stcgan_pack_weight_pad16")


def unpack_input_grad(g, coff, cn, grad_nchw, accumulate):
    n, h, w, c, ldg = _nhwc(g)
    assert grad_nchw.dtype == torch.float32 and grad_nchw.is_contiguous()
    check(_lib.load().stcgan_unpack_input_grad(_code(g), g.data_ptr(), n, h, w, ldg, coff, cn, grad_nchw.data_ptr(),
                                               int(accumulate), _stream()), "stcgan_unpack_input_grad")


def nhwc_to_nchw(x):
    n, h, w, c, ld = _nhwc(x)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().stcgan_nhwc_to_nchw(_code(x), x.data_ptr(), n, h, w, c, ld, out.data_ptr(), _stream()), "stcgan_nhwc_to_nchw")
    return out


def nchw_to_nhwc(x, dtype):
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), dtype=dtype, device=x.device)
    check(_lib.load().stcgan_nchw_to_nhwc(_code(out), x.data_ptr(), n, h, w, c, out.data_ptr(), c, _stream()), "stcgan_nchw_to_nhwc")
    return out


def out_act_bwd(act, out_nchw, dout_nchw, dtype, cpad=None, border=0):
    """gradient through the output activation, NCHW fp32 -> NHWC `dtype`; with `border`/`cpad` the result is the
    zero-bordered, channel-padded layout [N, H+2b, W+2b, cpad] the thin tensor-core kernels read."""
    n, c, h, w = out_nchw.shape
    assert dout_nchw.is_contiguous() and out_nchw.is_contiguous() and dout_nchw.dtype == torch.float32
    ld = c if cpad is None else cpad
    g = torch.empty((n, h + 2 * border, w + 2 * border, ld), dtype=dtype, device=out_nchw.device)
    check(_lib.load().stcgan_out_act_bwd(_code(g), act, out_nchw.data_ptr(), dout_nchw.data_ptr(), n, h, w, c, border,
                                         g.data_ptr(), ld, _stream()), "stcgan_out_act_bwd")
    return g


# single-launch BatchNorm backward for tiny tensors (stcgan_bn_act_bwd_small): OPT-IN (STCGAN_SMALL_BN=1).  Measured on B200 it
# does not pay: the four 64-pixel BatchNorm layers of a train step took 6 us longer each in one single-block launch than in the
# two-pass form (6.16-6.19 against 6.12 ms per step) -- one SM cannot hide three dependent global round trips any better than
# two small grids do.
SMALL_BN_ELEMS = 64 * 512
_SMALL_BN = _os.environ.get("STCGAN_SMALL_BN", "0") == "1"

KIND_L1, KIND_MSE, KIND_BCE = 0, 1, 2


def fused_loss(terms, loss_out):
    """terms: list of dict(kind, a, b=None, grad=None, target=0., weight=1., slot=0, accumulate=False)."""
    arr = (LossTerm * len(terms))()
    for i, t in enumerate(terms):
        a = t["a"]
        _need_cuda(a)
        assert a.dtype == torch.float32 and a.is_contiguous()
        arr[i].a = a.data_ptr()
        b = t.get("b")
        if b is not None:
            assert b.dtype == torch.float32 and b.is_contiguous() and b.numel() == a.numel()
        arr[i].b = None if b is None else b.data_ptr()
        g = t.get("grad")
        if g is not None:
            assert g.dtype == torch.float32 and g.is_contiguous() and g.numel() == a.numel()
        arr[i].grad = None if g is None else g.data_ptr()
        arr[i].n = a.numel()
        arr[i].target = float(t.get("target", 0.0))
        arr[i].weight = float(t.get("weight", 1.0))
        arr[i].loss_weight = float(t.get("loss_weight", t.get("weight", 1.0)))
        arr[i].kind = int(t["kind"])
        arr[i].slot = int(t.get("slot", 0))
        arr[i].accumulate = int(bool(t.get("accumulate", False)))
    check(_lib.load().stcgan_fused_loss(arr, len(terms), loss_out.data_ptr(), _stream()), "stcgan_fused_loss")


def rel_logits(a, b, avg, backward=False):
    """relativistic logits a - b (RpGAN) / a - mean_batch(b) (RaGAN), or with backward=True the gradient of that map
    w.r.t. b for the upstream gradient passed as `b` (a ignored)."""
    _need_cuda(b)
    assert b.dtype == torch.float32 and b.is_contiguous() and (backward or (a.dtype == torch.float32 and a.is_contiguous()
                                                                                and a.shape == b.shape))
    n = b.shape[0]
    mm = b.numel() // max(n, 1)
    out = torch.empty_like(b)
    check(_lib.load().stcgan_rel_logits(None if backward else a.data_ptr(), b.data_ptr(), n, mm, int(avg), int(backward),
                                        out.data_ptr(), _stream()), "stcgan_rel_logits")
    return out


def float2uint_hwc(x_nchw):
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    n, c, h, w = x_nchw.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x_nchw.device)
    check(_lib.load().stcgan_float2uint_hwc(x_nchw.data_ptr(), n, c, h, w, out.data_ptr(), _stream()), "stcgan_float2uint_hwc")
    return out


def u8_to_nchw(img_u8, out=None):
    """uint8 [N,H,W,C] (decoded images as cv2 gives them) -> float32 [N,C,H,W] in [-1,1]: the dataset transform on the GPU."""
    _need_cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and img_u8.dim() == 4
    n, h, w, c = img_u8.shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=img_u8.device)
    check(_lib.load().stcgan_u8_hwc_to_nchw_f32(img_u8.data_ptr(), n, h, w, c, out.data_ptr(), _stream()), "stcgan_u8_hwc_to_nchw_f32")
    return out


def float2uint(x):
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_lib.load().stcgan_float2uint(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "stcgan_float2uint")
    return out
