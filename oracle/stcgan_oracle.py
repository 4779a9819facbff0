"""CPU oracle for the ST-CGAN hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with plain functional torch ops on the CPU, the arithmetic
that the reference executes for the path named in BASELINE.json `north_star`.
It is imported only by `tests/`, by `__graft_entry__.smoke()` and by the
`cpu_baseline` / `--impl reference` legs of `bench.py`; the product package
(`shadow-removal-istd_b200/stcgan_b200`) never imports it.

Pinning status: the reference ships NO tests, golden vectors or saved weights
(SURVEY.md §4, §8c), so parity is pinned against *outputs of the reference
itself run in the build container*:
  * `oracle/pin_against_reference.py` imports the unmodified modules from
    /root/reference and checks every function below against them bit-for-bit;
  * `tests/golden/make_golden.py` (committed) ran the reference modules and
    wrote `tests/golden/*.npz`; `tests/test_cpu.py` re-checks this
    oracle against those fixtures everywhere (the GPU box has no /root/reference).

Reference lines each function follows are cited in its docstring
(paths relative to /root/reference/).

The third-party arithmetic provider of the reference is PyTorch (pinned
torch==1.5.1, environment.yml:24,159); this container has torch 2.11 whose
definitions of conv2d / conv_transpose2d / batch_norm / leaky_relu / relu /
tanh / l1 / mse / bce-with-logits / Adam are unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default (src/models/stcgan_g.py:15)
BN_MOMENTUM = 0.1
LEAK = 0.2             # nn.LeakyReLU(0.2, True) (stcgan_g.py:87, stcgan_d.py:24)
REFERENCE_SEED = 38107943   # src/main.py:239


# --------------------------------------------------------------------------
# state-dict construction (same RNG consumption order as the reference ctors)
# --------------------------------------------------------------------------

def generator_level_channels(in_channels, out_channels, ngf=64, num_downs=8):
    """Per-level (down_in, down_out, up_in, up_out), level 1 = outermost.

    Follows UnetGenerator.__init__ (src/models/stcgan_g.py:32-53): innermost
    block ngf*8 -> ngf*8, `num_downs-5` blocks at ngf*8, then ngf*4, ngf*2, ngf,
    and the outermost block from `in_channels` to `out_channels`.
    """
    widths = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)
    levels = []
    for k in range(1, num_downs + 1):
        d_in = in_channels if k == 1 else widths[k - 2]
        d_out = widths[k - 1]
        u_in = d_out if k == num_downs else 2 * d_out
        u_out = out_channels if k == 1 else widths[k - 2]
        levels.append((d_in, d_out, u_in, u_out))
    return levels


def generator_prefix(level):
    """state-dict prefix of the Sequential at U-Net `level` (1 = outermost).

    UnetGenerator.model is the outermost block, whose `.model` Sequential holds
    the next block at index 1 (outermost, stcgan_g.py:96-98) or 3
    (intermediate, stcgan_g.py:110-116).
    """
    p = "model.model."
    for k in range(1, level):
        p += ("1." if k == 1 else "3.") + "model."
    return p


def generator_keys(level, num_downs=8):
    """Names of (downconv, downnorm, upconv, upnorm) inside a level's Sequential."""
    p = generator_prefix(level)
    if level == 1:            # [downconv, sub, uprelu, upconv, tanh]
        return p + "0", None, p + "3", None
    if level == num_downs:    # [downrelu, downconv, uprelu, upconv, upnorm]
        return p + "1", None, p + "3", p + "4"
    return p + "1", p + "2", p + "5", p + "6"   # [lrelu, conv, bn, sub, relu, convT, bn]


def _bn_entries(sd, name, ch):
    bn = nn.BatchNorm2d(ch)
    for k, v in bn.state_dict().items():
        sd[f"{name}.{k}"] = v.clone()


def build_generator_state(in_channels, out_channels, ngf=64, num_downs=8):
    """Default-initialised state dict, drawing from torch's global RNG in the
    order the reference constructor does (innermost block first; inside a block:
    downconv, then the norms, then upconv -- stcgan_g.py:85-109)."""
    lv = generator_level_channels(in_channels, out_channels, ngf, num_downs)
    tmp = {}
    for level in range(num_downs, 0, -1):
        d_in, d_out, u_in, u_out = lv[level - 1]
        kd, kdn, ku, kun = generator_keys(level, num_downs)
        down = nn.Conv2d(d_in, d_out, 4, 2, 1, bias=False)
        up = nn.ConvTranspose2d(u_in, u_out, 4, 2, 1, bias=(level == 1))
        tmp[kd + ".weight"] = down.weight.detach().clone()
        if kdn:
            _bn_entries(tmp, kdn, d_out)
        tmp[ku + ".weight"] = up.weight.detach().clone()
        if level == 1:
            tmp[ku + ".bias"] = up.bias.detach().clone()
        if kun:
            _bn_entries(tmp, kun, u_out)
    # state_dict order = module registration order = outermost first.
    sd = OrderedDict()
    for level in range(1, num_downs + 1):
        kd, kdn, _, _ = generator_keys(level, num_downs)
        sd[kd + ".weight"] = tmp[kd + ".weight"]
        if kdn:
            for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                sd[f"{kdn}.{s}"] = tmp[f"{kdn}.{s}"]
    for level in range(num_downs, 0, -1):
        _, _, ku, kun = generator_keys(level, num_downs)
        sd[ku + ".weight"] = tmp[ku + ".weight"]
        if level == 1:
            sd[ku + ".bias"] = tmp[ku + ".bias"]
        if kun:
            for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                sd[f"{kun}.{s}"] = tmp[f"{kun}.{s}"]
    return sd


def discriminator_layers(in_channels, ndf=64, n_layers=3):
    """[(seq_index, cin, cout, stride, has_bias, bn_index|None)] per
    NLayerDiscriminator.__init__ (src/models/stcgan_d.py:21-50)."""
    layers = [(0, in_channels, ndf, 2, True, None)]
    idx, mult = 2, 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        layers.append((idx, ndf * prev, ndf * mult, 2, False, idx + 1))
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    layers.append((idx, ndf * prev, ndf * mult, 1, False, idx + 1))
    idx += 3
    layers.append((idx, ndf * mult, 1, 1, True, None))
    return layers


def build_discriminator_state(in_channels, ndf=64, n_layers=3):
    sd = OrderedDict()
    for (i, cin, cout, stride, bias, bn_i) in discriminator_layers(in_channels, ndf, n_layers):
        conv = nn.Conv2d(cin, cout, 4, stride, 1, bias=bias)
        sd[f"model.{i}.weight"] = conv.weight.detach().clone()
        if bias:
            sd[f"model.{i}.bias"] = conv.bias.detach().clone()
        if bn_i is not None:
            _bn_entries(sd, f"model.{bn_i}", cout)
    return sd


def build_all_states(seed=REFERENCE_SEED, ngf=64, ndf=64):
    """G1, G2, D1, D2 in the order src/cgan.py:35-66 constructs them."""
    torch.manual_seed(seed)
    return OrderedDict(
        G1=build_generator_state(3, 1, ngf),
        G2=build_generator_state(4, 3, ngf),
        D1=build_discriminator_state(4, ndf),
        D2=build_discriminator_state(7, ndf),
    )


def apply_weights_init(sd, gen=None):
    """`weights_init` (src/networks.py:19-30): conv AND BatchNorm weights ~ N(0, 0.02),
    biases 0.  Applied in `module.apply` order (children before parents, registration
    order), which for these nets equals state-dict order of the parameter holders."""
    for k in list(sd.keys()):
        if k.endswith(".weight"):
            sd[k] = torch.empty_like(sd[k]).normal_(0.0, 0.02, generator=gen)
            b = k[:-6] + "bias"
            if b in sd:
                sd[b] = torch.zeros_like(sd[b])
    return sd


def trainable_keys(sd):
    """Keys that are nn.Parameters, in `named_parameters()` order."""
    return [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


# --------------------------------------------------------------------------
# optional reduced-precision emulation of the convolutions (test infrastructure for the bf16 parity bounds)
# --------------------------------------------------------------------------
_EMULATE = None     # None = exact arithmetic of the tensors' dtype; torch.bfloat16 = see `emulate_conv_precision`


class emulate_conv_precision:
    """Context manager: every convolution rounds its input, its weight, its OUTPUT and the incoming output-gradient to
    `dtype` (values stay stored in the surrounding float32 / float64 tensors; products and sums are exact in that wider
    type).  This is the arithmetic model of a bf16 tensor-core pipeline with bf16 activation storage and fp32 accumulation
    and parameters; running the otherwise unchanged restatement under it gives the error that rounding alone causes
    (14-31 % on end-to-end gradients through gate flips), against which a CUDA bf16 path is bounded."""

    def __init__(self, dtype=torch.bfloat16):
        self.dtype = dtype

    def __enter__(self):
        global _EMULATE
        self.prev, _EMULATE = _EMULATE, self.dtype

    def __exit__(self, *a):
        global _EMULATE
        _EMULATE = self.prev


class _RoundSTE(torch.autograd.Function):
    """forward: round to `dtype`; backward: identity (the rounding of gradients is explicit, see _RoundGrad)."""

    @staticmethod
    def forward(ctx, x, dtype):
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _RoundGrad(torch.autograd.Function):
    """forward: identity; backward: round the gradient to `dtype`."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


def _conv(fn, x, w, bias, stride, pad, last=False):
    """`last`: the network's output layer, whose result (and incoming gradient) stay in float32 in the modelled pipeline."""
    if _EMULATE is None:
        return fn(x, w, bias, stride, pad)
    y = fn(_RoundSTE.apply(x, _EMULATE), _RoundSTE.apply(w, _EMULATE), None, stride, pad)
    if not last:
        y = _RoundGrad.apply(y, _EMULATE)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y if last else _RoundSTE.apply(y, _EMULATE)


# --------------------------------------------------------------------------
# forward passes (functional; autograd tracks through `sd` tensors)
# --------------------------------------------------------------------------

def _bn(sd, name, x, training):
    """nn.BatchNorm2d forward incl. running-stat update (stcgan_g.py:88,90)."""
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    if training:
        sd[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"],
                        training, BN_MOMENTUM, BN_EPS)


def generator_forward(sd, x, training=True, num_downs=8):
    """UnetGenerator.forward (stcgan_g.py:55-57, 120-132), unrolled.

    In-place LeakyReLU on the skip followed by the parent's in-place ReLU on the
    concat leaves relu(skip) in the decoder input; odd H/W are
    zero-padded bottom/right before a block and cropped after (stcgan_g.py:126-132).
    """
    skips, crops = [], []
    h = x
    for level in range(1, num_downs + 1):
        kd, kdn, _, _ = generator_keys(level, num_downs)
        if level > 1:
            skips.append(h)                       # block input = skip half
            ph, pw = h.size(2) % 2, h.size(3) % 2
            crops.append((h.size(2), h.size(3)))
            if ph or pw:
                h = F.pad(h, (0, pw, 0, ph))
            h = F.leaky_relu(h, LEAK)
        h = _conv(F.conv2d, h, sd[kd + ".weight"], None, 2, 1)
        if kdn:
            h = _bn(sd, kdn, h, training)
    for level in range(num_downs, 0, -1):
        _, _, ku, kun = generator_keys(level, num_downs)
        h = F.relu(h)
        h = _conv(F.conv_transpose2d, h, sd[ku + ".weight"], sd.get(ku + ".bias"), 2, 1, last=level == 1)
        if level == 1:
            return torch.tanh(h)
        h = _bn(sd, kun, h, training)
        hh, ww = crops.pop()
        h = torch.cat([skips.pop(), h[:, :, :hh, :ww]], 1)


def discriminator_forward(sd, x, training=True, n_layers=3, use_sigmoid=False, ndf=None):
    """NLayerDiscriminator.forward (stcgan_d.py:57-58)."""
    if ndf is None:
        ndf = sd["model.0.weight"].shape[0]
    layers = discriminator_layers(x.size(1), ndf, n_layers)
    h = x
    for n, (i, cin, cout, stride, bias, bn_i) in enumerate(layers):
        h = _conv(F.conv2d, h, sd[f"model.{i}.weight"], sd.get(f"model.{i}.bias"), stride, 1, last=n == len(layers) - 1)
        if bn_i is not None:
            h = _bn(sd, f"model.{bn_i}", h, training)
        if n != len(layers) - 1:
            h = F.leaky_relu(h, LEAK)
    return torch.sigmoid(h) if use_sigmoid else h


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------

def cal_loss(c, label, ls=False):
    """AdversarialLoss.cal_loss (src/loss.py:79-84) -- note the inverted flag:
    ls=False -> MSE, ls=True -> BCE-with-logits."""
    t = torch.as_tensor(label, dtype=c.dtype, device=c.device).expand_as(c)
    return F.mse_loss(c, t) if not ls else F.binary_cross_entropy_with_logits(c, t)


def adversarial_loss(c_real, c_fake, d_loss=True, ls=False, rel=False, avg=False):
    """AdversarialLoss.forward (src/loss.py:86-112); labels per loss.py:68-74."""
    real, fake = 1.0, (-1.0 if ls else 0.0)
    if d_loss:
        if rel and avg:
            return 0.5 * (cal_loss(c_real - c_fake.mean(dim=0), real, ls)
                          + cal_loss(c_fake - c_real.mean(dim=0), fake, ls))
        if rel:
            return cal_loss(c_real - c_fake, real, ls)
        return 0.5 * (cal_loss(c_real, real, ls) + cal_loss(c_fake, fake, ls))
    if rel and avg:
        return 0.5 * (cal_loss(c_real - c_fake.mean(dim=0), fake, ls)
                      + cal_loss(c_fake - c_real.mean(dim=0), real, ls))
    if rel:
        return cal_loss(c_fake - c_real, real, ls)
    return cal_loss(c_fake, real, ls)


def data_loss(pred, target):
    """DataLoss.forward (src/loss.py:25-26): L1, mean reduction."""
    return F.l1_loss(pred, target, reduction="mean")


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------

def make_istd_batch(batch, height=256, width=256, seed=42, binary_matte=False):
    """Synthetic ISTD-shaped triple on the dataset's value grid
    (src/dataset.py:100-110,152): uint8 -> /255 -> (v-0.5)*2, CHW float32."""
    rng = np.random.RandomState(seed)
    lo_h, lo_w = max(height // 8, 1), max(width // 8, 1)

    def smooth(c):
        z = rng.rand(batch, lo_h, lo_w, c).astype(np.float32)
        z = np.repeat(np.repeat(z, -(-height // lo_h), 1), -(-width // lo_w), 2)[:, :height, :width]
        z = 0.7 * z + 0.3 * rng.rand(batch, height, width, c).astype(np.float32)
        return z

    img = np.clip(smooth(3) * 255, 0, 255).astype(np.uint8)
    blob = smooth(1)
    if binary_matte:
        matte = np.where(blob > 0.55, 255, 0).astype(np.uint8)
    else:
        matte = np.clip((blob - 0.35) * 3.0, 0, 1)
        matte = (matte * 255).astype(np.uint8)
    gain = 1.0 + 0.6 * (matte.astype(np.float32) / 255.0)
    target = np.clip(img.astype(np.float32) * gain, 0, 255).astype(np.uint8)

    def to_tensor(u8):
        f = u8.astype(np.float32) / 255                  # utils.uint2float (src/utils.py:60-62)
        return torch.as_tensor((f.transpose(0, 3, 1, 2) - 0.5) * 2, dtype=torch.float32)

    return to_tensor(img), to_tensor(matte), to_tensor(target)


# --------------------------------------------------------------------------
# the train step and inference, restated
# --------------------------------------------------------------------------

@dataclass
class HyperParams:
    """Defaults of src/main.py:182-239 (VisualLoss weights forced to 0)."""
    lr_G: float = 5e-4
    lr_D: float = 1e-4
    beta1: float = 0.5
    beta2: float = 0.999
    lambda1: float = 5.0
    lambda2: float = 0.5
    lambda3: float = 0.5
    lambda4: float = 0.0    # weights of vis1 / vis2 (cgan.py:347-348; reference default 5 / 50, needs VGG19 weights)
    lambda5: float = 0.0
    ls: bool = False        # what `args.D_loss_fn == "leastsqure"` evaluates to (cgan.py:147)
    rel: bool = False
    avg: bool = False


class OracleTrainer:
    """Holds G1/G2/D1/D2 state dicts + two Adam optimisers (cgan.py:85-90) and
    runs the step body of CGAN.run_epoch (src/cgan.py:274-351) with
    lambda4 = lambda5 = 0."""

    def __init__(self, states, hp: HyperParams = HyperParams(), dtype=torch.float32, visual_loss=None):
        self.hp = hp
        self.visual_loss = visual_loss     # callable (pred, target) -> scalar standing in for VisualLoss (loss.py:29-56)
        self.sd = OrderedDict()
        for name, sd in states.items():
            self.sd[name] = OrderedDict(
                (k, (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()))
                for k, v in sd.items())
        self.params = {}
        for name, sd in self.sd.items():
            ps = []
            for k in trainable_keys(sd):
                sd[k].requires_grad_(True)
                ps.append(sd[k])
            self.params[name] = ps
        self.optim_G = torch.optim.Adam(self.params["G1"] + self.params["G2"],
                                        lr=hp.lr_G, betas=(hp.beta1, hp.beta2))
        self.optim_D = torch.optim.Adam(self.params["D1"] + self.params["D2"],
                                        lr=hp.lr_D, betas=(hp.beta1, hp.beta2))

    def _req(self, names, flag):
        for n in names:
            for p in self.params[n]:
                p.requires_grad_(flag)

    def forward_only(self, x, m, y, training=True):
        """BASELINE configs[0]: G1+G2 forward, D1/D2 on real and fake, both losses."""
        hp, sd = self.hp, self.sd
        with torch.no_grad():
            out = {}
            out["C1_real"] = discriminator_forward(sd["D1"], torch.cat((x, m), 1), training)
            out["m_pred"] = generator_forward(sd["G1"], x, training)
            out["C1_fake"] = discriminator_forward(sd["D1"], torch.cat((x, out["m_pred"]), 1), training)
            out["C2_real"] = discriminator_forward(sd["D2"], torch.cat((x, m, y), 1), training)
            out["y_pred"] = generator_forward(sd["G2"], torch.cat((x, out["m_pred"]), 1), training)
            out["C2_fake"] = discriminator_forward(
                sd["D2"], torch.cat((x, out["m_pred"], out["y_pred"]), 1), training)
            adv = lambda r, f, d: adversarial_loss(r, f, d, hp.ls, hp.rel, hp.avg)
            out["D1_loss"] = adv(out["C1_real"], out["C1_fake"], True)
            out["D2_loss"] = adv(out["C2_real"], out["C2_fake"], True)
            out["D_loss"] = hp.lambda2 * out["D1_loss"] + hp.lambda3 * out["D2_loss"]
            out["G1_loss"] = adv(out["C1_real"], out["C1_fake"], False)
            out["G2_loss"] = adv(out["C2_real"], out["C2_fake"], False)
            out["data1_loss"] = data_loss(out["m_pred"], m)
            out["data2_loss"] = data_loss(out["y_pred"], y)
            out["G_loss"] = (out["data1_loss"] + hp.lambda1 * out["data2_loss"]
                             + hp.lambda2 * out["G1_loss"] + hp.lambda3 * out["G2_loss"])
        return out

    def train_step(self, x, m, y, do_optim=True, keep_grads=False):
        hp, sd = self.hp, self.sd
        adv = lambda r, f, d: adversarial_loss(r, f, d, hp.ls, hp.rel, hp.avg)
        out = {}
        self.optim_D.zero_grad(); self.optim_G.zero_grad()          # cgan.py:274-276
        self._req(("D1", "D2"), True)                                # cgan.py:278-279
        c1_real = discriminator_forward(sd["D1"], torch.cat((x, m), 1))              # :281
        m_pred = generator_forward(sd["G1"], x)                                      # :282
        c1_fake = discriminator_forward(sd["D1"], torch.cat((x, m_pred.detach()), 1))  # :283
        c2_real = discriminator_forward(sd["D2"], torch.cat((x, m, y), 1))           # :285
        y_pred = generator_forward(sd["G2"], torch.cat((x, m_pred), 1))              # :286
        c2_fake = discriminator_forward(
            sd["D2"], torch.cat((x, m_pred.detach(), y_pred.detach()), 1))           # :287-289
        d1 = adv(c1_real, c1_fake, True); d2 = adv(c2_real, c2_fake, True)           # :299-300
        d_loss = hp.lambda2 * d1 + hp.lambda3 * d2                                   # :302
        d_loss.backward()                                                            # :304
        if keep_grads:
            out["grads_D"] = {n: [p.grad.detach().clone() for p in self.params[n]] for n in ("D1", "D2")}
        if do_optim:
            self.optim_D.step()                                                      # :305
        out.update(C1_real_Dphase=c1_real.detach(), C1_fake_Dphase=c1_fake.detach(),
                   C2_real_Dphase=c2_real.detach(), C2_fake_Dphase=c2_fake.detach(),
                   D1_loss=d1.detach(), D2_loss=d2.detach(), D_loss=d_loss.detach())
        self.optim_G.zero_grad()                                                     # :316
        self._req(("D1", "D2"), False)                                               # :317-318
        c1_real = discriminator_forward(sd["D1"], torch.cat((x, m), 1))              # :321
        c1_fake = discriminator_forward(sd["D1"], torch.cat((x, m_pred), 1))         # :322
        c2_real = discriminator_forward(sd["D2"], torch.cat((x, m, y), 1))           # :323
        c2_fake = discriminator_forward(sd["D2"], torch.cat((x, m_pred, y_pred), 1))  # :324
        g1 = adv(c1_real, c1_fake, False); g2 = adv(c2_real, c2_fake, False)         # :329-330
        data1 = data_loss(m_pred, m); data2 = data_loss(y_pred, y)                   # :332-333
        g_loss = data1 + hp.lambda1 * data2 + hp.lambda2 * g1 + hp.lambda3 * g2      # :343-348
        if self.visual_loss is not None:                                             # :334-336, 347-348
            vis1 = self.visual_loss(m_pred.expand(-1, 3, -1, -1), m.expand(-1, 3, -1, -1))
            vis2 = self.visual_loss(y_pred, y)
            g_loss = g_loss + hp.lambda4 * vis1 + hp.lambda5 * vis2
            out.update(vis1_loss=vis1.detach(), vis2_loss=vis2.detach())
        g_loss.backward()                                                            # :350
        if keep_grads:
            out["grads_G"] = {n: [p.grad.detach().clone() for p in self.params[n]] for n in ("G1", "G2")}
        if do_optim:
            self.optim_G.step()                                                      # :351
        self._req(("D1", "D2"), True)
        out.update(m_pred=m_pred.detach(), y_pred=y_pred.detach(),
                   C1_fake_Gphase=c1_fake.detach(), C2_fake_Gphase=c2_fake.detach(),
                   G1_loss=g1.detach(), G2_loss=g2.detach(), data1_loss=data1.detach(),
                   data2_loss=data2.detach(), G_loss=g_loss.detach())
        return out


class OracleDataParallel:
    """Single-process statement of the data-parallel step (nn.DataParallel semantics of src/cgan.py:78-84):
    every shard goes through its OWN forward (per-replica BatchNorm batch statistics and running buffers), the loss of a
    shard is the mean over that shard, parameter gradients are the average over shards, ONE Adam update is applied to the
    shared parameters.  `shards` = list of (x, m, y).  Returns per-shard outputs of `OracleTrainer.train_step`'s keys plus
    the averaged gradients."""

    def __init__(self, states, world, hp: HyperParams = HyperParams(), dtype=torch.float32):
        self.t = OracleTrainer(states, hp, dtype)
        self.world, self.hp = world, hp
        # shard r sees the shared parameter tensors and its own copy of the BatchNorm buffers
        self.views = []
        for r in range(world):
            v = OrderedDict()
            for name, sd in self.t.sd.items():
                keys = set(trainable_keys(sd))
                v[name] = OrderedDict((k, (t_ if k in keys else t_.clone())) for k, t_ in sd.items())
            self.views.append(v)

    def train_step(self, shards, keep_grads=True):
        hp, t, W = self.hp, self.t, self.world
        adv = lambda r, f, d: adversarial_loss(r, f, d, hp.ls, hp.rel, hp.avg)
        outs = [dict() for _ in shards]
        t.optim_D.zero_grad(); t.optim_G.zero_grad()
        t._req(("D1", "D2"), True)
        held, total = [], 0.0
        for o, sd, (x, m, y) in zip(outs, self.views, shards):
            c1r = discriminator_forward(sd["D1"], torch.cat((x, m), 1))
            mp = generator_forward(sd["G1"], x)
            c1f = discriminator_forward(sd["D1"], torch.cat((x, mp.detach()), 1))
            c2r = discriminator_forward(sd["D2"], torch.cat((x, m, y), 1))
            yp = generator_forward(sd["G2"], torch.cat((x, mp), 1))
            c2f = discriminator_forward(sd["D2"], torch.cat((x, mp.detach(), yp.detach()), 1))
            d1, d2 = adv(c1r, c1f, True), adv(c2r, c2f, True)
            total = total + (hp.lambda2 * d1 + hp.lambda3 * d2) / W
            o.update(D1_loss=d1.detach(), D2_loss=d2.detach())
            held.append((mp, yp))
        total.backward()
        grads = {}
        if keep_grads:
            grads.update({n: [p.grad.detach().clone() for p in t.params[n]] for n in ("D1", "D2")})
        t.optim_D.step()
        t.optim_G.zero_grad()
        t._req(("D1", "D2"), False)
        total = 0.0
        for o, sd, (x, m, y), (mp, yp) in zip(outs, self.views, shards, held):
            c1r = discriminator_forward(sd["D1"], torch.cat((x, m), 1))
            c1f = discriminator_forward(sd["D1"], torch.cat((x, mp), 1))
            c2r = discriminator_forward(sd["D2"], torch.cat((x, m, y), 1))
            c2f = discriminator_forward(sd["D2"], torch.cat((x, mp, yp), 1))
            g1, g2 = adv(c1r, c1f, False), adv(c2r, c2f, False)
            da1, da2 = data_loss(mp, m), data_loss(yp, y)
            gl = da1 + hp.lambda1 * da2 + hp.lambda2 * g1 + hp.lambda3 * g2
            total = total + gl / W
            o.update(m_pred=mp.detach(), y_pred=yp.detach(), G1_loss=g1.detach(), G2_loss=g2.detach(),
                     data1_loss=da1.detach(), data2_loss=da2.detach(), G_loss=gl.detach())
        total.backward()
        if keep_grads:
            grads.update({n: [p.grad.detach().clone() for p in t.params[n]] for n in ("G1", "G2")})
        t.optim_G.step()
        t._req(("D1", "D2"), True)
        return outs, grads


def float2uint(array):
    """utils.float2uint (src/utils.py:65-67): clip to [0,1], *255, truncate."""
    assert array.dtype in (np.float32, np.float64)
    return (np.clip(array, 0, 1) * 255).astype(np.uint8)


def infer(sd_g1, sd_g2, x):
    """CGAN.infer core (src/cgan.py:437-442, 452-458): eval-mode G1 -> G2, then
    `*0.5+0.5` in numpy float32, HWC transpose, truncating uint8 quantisation."""
    with torch.no_grad():
        m_pred = generator_forward(sd_g1, x, training=False)
        y_pred = generator_forward(sd_g2, torch.cat((x, m_pred), 1), training=False)
    m_np = m_pred.numpy() * 0.5 + 0.5
    y_np = y_pred.numpy() * 0.5 + 0.5
    m_u8 = np.stack([float2uint(m_np[i].transpose(1, 2, 0)) for i in range(m_np.shape[0])])
    y_u8 = np.stack([float2uint(y_np[i].transpose(1, 2, 0)) for i in range(y_np.shape[0])])
    return m_pred, y_pred, m_u8, y_u8


# --------------------------------------------------------------------------
# algorithmic work, used by bench.py for the roofline
# --------------------------------------------------------------------------

def conv_flops_generator(in_ch, out_ch, h, w, ngf=64, num_downs=8):
    """Dense conv/convT FLOPs (2*MACs) of one G forward on one image, following the
    padded spatial chain of stcgan_g.py:124-132.  Returns (total, e1_flops)."""
    lv = generator_level_channels(in_ch, out_ch, ngf, num_downs)
    total, e1, inner = 0, 0, []
    hh, ww = h, w
    for k, (d_in, d_out, _, _) in enumerate(lv, 1):
        if k > 1:
            hh, ww = hh + hh % 2, ww + ww % 2          # F.pad to even
        hh, ww = hh // 2, ww // 2
        inner.append((hh, ww))
        f = 2 * 16 * d_in * d_out * hh * ww
        total += f
        if k == 1:
            e1 = f
    for k in range(num_downs, 0, -1):
        _, _, u_in, u_out = lv[k - 1]
        hi, wi = inner[k - 1]
        total += 2 * 16 * u_in * u_out * hi * wi        # 16*Hi*Wi*Cin*Cout non-zero MACs
    return total, e1


def conv_flops_discriminator(in_ch, h, w, ndf=64, n_layers=3):
    total, c1 = 0, 0
    hh, ww = h, w
    for n, (i, cin, cout, stride, bias, bn_i) in enumerate(discriminator_layers(in_ch, ndf, n_layers)):
        hh = (hh + 2 - 4) // stride + 1
        ww = (ww + 2 - 4) // stride + 1
        f = 2 * 16 * cin * cout * hh * ww
        total += f
        if n == 0:
            c1 = f
    return total, c1


def train_step_flops(h=256, w=256):
    """Per-image algorithmic FLOPs of the full train step (SURVEY §8d formula)."""
    g1, g1e1 = conv_flops_generator(3, 1, h, w)
    g2, _ = conv_flops_generator(4, 3, h, w)
    d1, d1c1 = conv_flops_discriminator(4, h, w)
    d2, d2c1 = conv_flops_discriminator(7, h, w)
    fwd = g1 + g2 + 4 * (d1 + d2)
    bwd_d = 2 * (2 * d1 - d1c1) + 2 * (2 * d2 - d2c1)
    bwd_g = d1 + d2 + 2 * g2 + 2 * g1 - g1e1
    return fwd + bwd_d + bwd_g


def inference_flops(h=480, w=640):
    return conv_flops_generator(3, 1, h, w)[0] + conv_flops_generator(4, 3, h, w)[0]
