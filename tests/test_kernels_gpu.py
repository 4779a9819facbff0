"""Per-kernel parity: every C-ABI kernel against the single torch op it replaces (CPU, float64 arbiter),
fed IDENTICAL inputs (layer-local checks are where 1e-3 / 2e-2 are enforceable).

fp32 mode (CUDA-core tiles)            : tolerance 1e-5 (north_star asks 1e-3)
bf16 mode (tcgen05 / CUDA-core, bf16 IO): inputs are rounded to bf16 first, so the only error left is fp32
                                         accumulation order + the bf16 rounding of the output: tol 6e-3
                                         (north_star asks 2e-2)
"""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 6e-3}


def _nhwc(t, dtype, dev):
    n, c, h, w = t.shape
    out = torch.empty((n, h, w, c), dtype=dtype, device=dev)      # canonical strides even for size-1 dims
    out.copy_(t.permute(0, 2, 3, 1))
    return out


def _nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def _round(t, mode):
    return t.to(torch.bfloat16).float() if mode == "bf16" else t


def _convop(kind, cin, cout, mode, dev, bias=False, seed=0):
    from stcgan_b200 import nets, ops
    g = torch.Generator().manual_seed(seed)
    shape = (cin, cout, 4, 4) if kind == "convT" else (cout, cin, 4, 4)
    w = torch.randn(shape, generator=g) * (1.0 / (16 * cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1 if bias else None
    wp = torch.nn.Parameter(w.to(dev))
    bp = None if b is None else torch.nn.Parameter(b.to(dev))
    op = nets.ConvOp(kind, wp, bp)
    op.ensure_packed(ops.DTYPES[mode][1])
    op.g = torch.zeros(16 * op.d0 * op.d1, dtype=torch.float32, device=dev)
    op.gb = None if b is None else torch.zeros(cout, dtype=torch.float32, device=dev)
    return op, w, b


def _ref_forward(kind, x, w, b):
    if kind == "conv2":
        return F.conv2d(x, w, b, 2, 1)
    if kind == "conv1":
        return F.conv2d(x, w, b, 1, 1)
    return F.conv_transpose2d(x, w, b, 2, 1)


CASES = [
    # kind, cin, cout, n, h, w
    ("conv2", 64, 128, 2, 32, 32),     # e2-like (TC eligible)
    ("conv2", 128, 64, 3, 16, 24),     # BN=64 tile, non-square
    ("conv2", 64, 64, 2, 18, 14),      # small / ragged tiles
    ("conv2", 512, 512, 4, 2, 2),      # e8-like: 1x1 output
    ("conv2", 3, 64, 2, 32, 32),       # e1-like thin K (CUDA-core path in both modes)
    ("conv2", 7, 64, 1, 20, 28),       # D c1-like
    ("conv1", 256, 512, 2, 8, 8),      # D c4-like stride 1 (7x7 out)
    ("conv1", 128, 128, 1, 32, 32),    # 31x31 out like D c4 at 256^2
    ("conv1", 512, 1, 2, 9, 9),        # D c5-like thin N
    ("convT", 128, 64, 2, 16, 16),     # decoder-like
    ("convT", 1024, 512, 2, 2, 2),     # d7-like
    ("convT", 512, 512, 4, 1, 1),      # d8-like: 1x1 input
    ("convT", 128, 3, 2, 16, 16),      # d1-like thin N
    ("convT", 64, 128, 1, 5, 7),       # odd input grid
]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_forward_dgrad_wgrad(cuda, lib, mode, case):
    from stcgan_b200 import ops
    kind, cin, cout, n, h, w_ = case
    dt = ops.DTYPES[mode][1]
    op, w, b = _convop(kind, cin, cout, mode, cuda, bias=(cout <= 64))
    g = torch.Generator().manual_seed(1)
    x = _round(torch.randn(n, cin, h, w_, generator=g), mode)
    wr = _round(w, mode)
    xd = x.double().requires_grad_(True)
    wd = wr.double().requires_grad_(True)
    ref = _ref_forward(kind, xd, wd, None if b is None else b.double())
    oh, ow = ref.shape[2:]
    assert (oh, ow) == op.out_size(h, w_)
    out = op.forward(_nhwc(x, dt, cuda), oh, ow)
    torch.cuda.synchronize()
    assert rel_err(_nchw(out), ref) < TOL[mode], "forward"
    # backward with a random output gradient
    go = _round(torch.randn(ref.shape, generator=g), mode)
    ref.backward(go.double())
    gx = op.dgrad(_nhwc(go, dt, cuda), h, w_)
    torch.cuda.synchronize()
    assert rel_err(_nchw(gx), xd.grad) < TOL[mode], "dgrad"
    op.wgrad(_nhwc(x, dt, cuda), _nhwc(go, dt, cuda))
    gw = ops.unpack_grad(op.g, op.d0, op.d1)
    torch.cuda.synchronize()
    assert rel_err(gw, wd.grad) < (1e-5 if mode == "fp32" else 2e-5), "wgrad (fp32 accumulation of exact bf16 products)"
    if b is not None:
        assert rel_err(op.gb, go.double().sum(dim=(0, 2, 3))) < 1e-5, "bias grad"


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_reads_padding_out_of_range_and_writes_channel_slices(cuda, lib, mode):
    """odd-size F.pad (stcgan_g.py:126-132) as out-of-range reads; in-place concat via channel-slice outputs."""
    from stcgan_b200 import ops
    dt = ops.DTYPES[mode][1]
    op, w, _ = _convop("conv2", 64, 64, mode, cuda)
    g = torch.Generator().manual_seed(3)
    x = _round(torch.randn(2, 64, 15, 5, generator=g), mode)        # 15x5 -> padded 16x6 -> 8x3
    ref = F.conv2d(F.pad(x.double(), (0, 1, 0, 1)), _round(w, mode).double(), None, 2, 1)
    buf = torch.zeros(2, 8, 3, 192, dtype=dt, device=cuda)
    op.forward(_nhwc(x, dt, cuda), 8, 3, out=buf[..., 64:128])
    torch.cuda.synchronize()
    assert rel_err(_nchw(buf[..., 64:128]), ref) < TOL[mode]
    assert float(buf[..., :64].abs().max()) == 0 and float(buf[..., 128:].abs().max()) == 0


BNSTAT_CASES = [
    ("conv2", 64, 128, 4, 32, 32),     # BN=128 tiles, several tiles per channel
    ("conv2", 128, 64, 3, 16, 24),     # BN=64 tiles
    ("conv2", 64, 64, 2, 15, 5),       # ragged tiles + out-of-range rows (must not count)
    ("conv1", 256, 512, 2, 8, 8),      # stride 1, 7x7 out
    ("convT", 128, 64, 2, 16, 16),     # four parity classes accumulate into the same channels
    ("convT", 64, 128, 1, 5, 7),       # odd grid
    ("convT", 1024, 512, 2, 2, 2),     # split-K path: statistics from the finisher kernel
    ("conv2", 512, 512, 4, 4, 4),      # split-K path, Conv2d
]


@pytest.mark.parametrize("case", BNSTAT_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_epilogue_batchnorm_statistics(cuda, lib, case):
    """BatchNorm batch statistics fused into the GEMM epilogue (stcgan_tapconv_bnstats) == sums over the bf16 conv output
    the same launch wrote; then finalize+apply in one launch (stcgan_bn_fused_apply) against nn.BatchNorm2d semantics."""
    from stcgan_b200 import _lib, ops
    from stcgan_b200._lib import ACT_LEAKY, ACT_RELU
    kind, cin, cout, n, h, w_ = case
    op, w, _ = _convop(kind, cin, cout, "bf16", cuda)
    assert op.stats_fusable()
    g = torch.Generator().manual_seed(5)
    x = _round(torch.randn(n, cin, h, w_, generator=g), "bf16")
    ph, pw = (h + h % 2, w_ + w_ % 2) if kind == "conv2" else (h, w_)
    oh, ow = op.out_size(ph, pw)
    acc = torch.zeros((_lib.BN_SLOTS, 2, cout), dtype=torch.float64, device=cuda)
    y = op.forward(_nhwc(x, torch.bfloat16, cuda), oh, ow, bn_acc=acc)
    y_plain = op.forward(_nhwc(x, torch.bfloat16, cuda), oh, ow)
    torch.cuda.synchronize()
    # (split-K layers reduce their partial sums with fp32 atomics in arbitrary order: equal up to a bf16 ulp only)
    assert rel_err(y.float(), y_plain.float().cpu().double()) < 3e-3, "the statistics epilogue must not change the conv output"
    yd = y.double().cpu().reshape(-1, cout)
    tot = acc.sum(dim=0).cpu()
    cnt = yd.shape[0]
    assert rel_err(tot[0], yd.sum(0)) < 1e-5 and rel_err(tot[1], (yd * yd).sum(0)) < 1e-5
    # finalize + apply, training mode, against torch on the same bf16 tensor
    gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    rm, rv = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
    mi = torch.empty((2, cout), device=cuda); ss = torch.empty((2, cout), device=cuda)
    o1, o2 = torch.empty_like(y), torch.empty_like(y)
    ops.bn_fused_apply(y, acc, cnt, gamma, beta, rm, rv, 0.1, 1e-5, True, mi, ss, o1, ACT_LEAKY, o2, ACT_RELU)
    torch.cuda.synchronize()
    bn = torch.nn.BatchNorm2d(cout).double()
    with torch.no_grad():
        bn.weight.copy_(gamma.double().cpu()); bn.bias.copy_(beta.double().cpu())
    z = bn(_nchw(y).double())
    assert rel_err(_nchw(o1), F.leaky_relu(z, 0.2)) < 6e-3 and rel_err(_nchw(o2), F.relu(z)) < 6e-3
    assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5
    assert rel_err(mi[0], yd.mean(0)) < 1e-5 and rel_err(mi[1], 1.0 / torch.sqrt(yd.var(0, unbiased=False) + 1e-5)) < 1e-5
    # eval mode: running statistics, no accumulator
    o3 = torch.empty_like(y)
    ops.bn_fused_apply(y, None, cnt, gamma, beta, rm, rv, 0.0, 1e-5, False, mi, ss, o3, ACT_RELU)
    torch.cuda.synchronize()
    bn.eval()
    assert rel_err(_nchw(o3), F.relu(bn(_nchw(y).double()))) < 6e-3


MT2_CASES = [
    ("conv2", 64, 128, 6, 128, 128),    # 192 tiles of 128 pixels -> 96 CTAs with two accumulators each
    ("convT", 128, 128, 6, 31, 33),     # four parity classes, odd grid, ragged 256-pixel tiles
    ("conv1", 128, 256, 8, 40, 40),     # stride 1, two N tiles
]


@pytest.mark.parametrize("case", MT2_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_two_accumulator_tiles(cuda, lib, case, monkeypatch):
    """M tiles of 2 x 128 pixels (two TMEM accumulators per CTA, STCGAN_TC_MT=2 forces the path the library picks on its own
    for multi-wave launches): forward + fused BatchNorm statistics against torch, and == the one-accumulator kernel."""
    from stcgan_b200 import _lib, ops
    kind, cin, cout, n, h, w_ = case
    op, w, _ = _convop(kind, cin, cout, "bf16", cuda)
    g = torch.Generator().manual_seed(7)
    x = _round(torch.randn(n, cin, h, w_, generator=g), "bf16")
    ref = _ref_forward(kind, x.double(), _round(w, "bf16").double(), None)
    oh, ow = ref.shape[2:]
    xk = _nhwc(x, torch.bfloat16, cuda)
    monkeypatch.setenv("STCGAN_TC_MT", "1")
    y1 = op.forward(xk, oh, ow)
    monkeypatch.setenv("STCGAN_TC_MT", "2")
    acc = torch.zeros((_lib.BN_SLOTS, 2, cout), dtype=torch.float64, device=cuda)
    y2 = op.forward(xk, oh, ow, bn_acc=acc)
    torch.cuda.synchronize()
    assert rel_err(_nchw(y2), ref) < TOL["bf16"]
    assert torch.equal(y1, y2), "same MMAs in the same order: bit-identical to the one-accumulator kernel"
    yd = y2.double().cpu().reshape(-1, cout)
    tot = acc.sum(dim=0).cpu()
    assert rel_err(tot[0], yd.sum(0)) < 1e-5 and rel_err(tot[1], (yd * yd).sum(0)) < 1e-5


def test_pack_weight_layout(cuda, lib):
    from stcgan_b200 import ops
    w = torch.randn(24, 40, 4, 4)
    for dt in (torch.float32, torch.bfloat16):
        p1 = torch.empty(16, 24, 40, dtype=dt, device=cuda); p2 = torch.empty(16, 40, 24, dtype=dt, device=cuda)
        ops.pack_weight(w.to(cuda), p1, p2)
        r1 = w.permute(2, 3, 0, 1).reshape(16, 24, 40).to(dt); r2 = w.permute(2, 3, 1, 0).reshape(16, 40, 24).to(dt)
        assert torch.equal(p1.cpu(), r1) and torch.equal(p2.cpu(), r2)
    g = torch.randn(16, 24, 40, device=cuda)
    assert torch.equal(ops.unpack_grad(g.reshape(-1), 24, 40).cpu(), g.cpu().permute(1, 2, 0).reshape(24, 40, 4, 4))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 9, 7, 64), (16, 4, 4, 512), (16, 2, 2, 512), (1, 2, 2, 8), (3, 31, 31, 128)])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_act_forward_backward(cuda, lib, mode, shape, training):
    """BN(train/eval) + LeakyReLU and ReLU dual output, with a crop, against torch ops in float64."""
    from stcgan_b200 import nets, ops
    from stcgan_b200._lib import ACT_LEAKY, ACT_RELU
    dt = ops.DTYPES[mode][1]
    n, h, w, c = shape
    hc, wc = max(h - 1, 1), max(w - 1, 1)
    g = torch.Generator().manual_seed(5)
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(c, generator=g) * 0.5 + 1); bn.bias.copy_(torch.randn(c, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(c, generator=g) * 0.1); bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    ref_bn = torch.nn.BatchNorm2d(c).double(); ref_bn.load_state_dict(bn.state_dict()); ref_bn.train(training)
    bn = bn.to(cuda).train(training)
    y = _round(torch.randn(n, c, h, w, generator=g) * 2 + 0.5, mode)
    yd = y.double().requires_grad_(True)
    z = ref_bn(yd)[:, :, :hc, :wc]
    r1, r2 = F.leaky_relu(z, 0.2), F.relu(z)
    op = nets.BNOp(bn)
    op.ggamma = torch.zeros(c, device=cuda); op.gbeta = torch.zeros(c, device=cuda)
    sc = nets._bn_scratch(c, cuda)
    yk = _nhwc(y, dt, cuda)
    o1 = torch.empty(n, hc, wc, c, dtype=dt, device=cuda); o2 = torch.empty_like(o1)
    op.forward(yk, sc, training, o1, ACT_LEAKY, o2, ACT_RELU)
    torch.cuda.synchronize()
    tol = 1e-5 if mode == "fp32" else 6e-3
    assert rel_err(_nchw(o1), r1) < tol and rel_err(_nchw(o2), r2) < tol
    if training:
        assert rel_err(bn.running_mean, ref_bn.running_mean) < 1e-5 and rel_err(bn.running_var, ref_bn.running_var) < 1e-5
        assert int(bn.num_batches_tracked) == 1
    g1 = _round(torch.randn(r1.shape, generator=g), mode); g2 = _round(torch.randn(r2.shape, generator=g), mode)
    (r1 * g1.double()).sum().backward(retain_graph=True); (r2 * g2.double()).sum().backward()
    dy = torch.empty_like(yk)
    op.backward(yk, sc, training, _nhwc(g1, dt, cuda), ACT_LEAKY, _nhwc(g2, dt, cuda), ACT_RELU, dy, True)
    torch.cuda.synchronize()
    # bf16 inputs near a gate (|z| ~ rounding) may flip it: compare with a loose norm-wise tolerance in bf16
    assert rel_err(_nchw(dy), yd.grad) < (1e-4 if mode == "fp32" else 2e-2)
    assert rel_err(op.ggamma, ref_bn.weight.grad) < (1e-4 if mode == "fp32" else 2e-2)
    assert rel_err(op.gbeta, ref_bn.bias.grad) < (1e-4 if mode == "fp32" else 2e-2)


def test_batchnorm_backward_single_launch_variant(cuda, lib, monkeypatch):
    """the opt-in single-launch backward for tiny tensors (stcgan_bn_act_bwd_small) against the same torch reference"""
    from stcgan_b200 import ops
    monkeypatch.setattr(ops, "_SMALL_BN", True)
    test_batchnorm_act_forward_backward(cuda, lib, "bf16", (16, 2, 2, 512), True)
    test_batchnorm_act_forward_backward(cuda, lib, "bf16", (2, 3, 5, 64), True)


def test_batchnorm_single_value_raises(cuda, lib):
    from stcgan_b200 import nets
    from stcgan_b200._lib import ACT_RELU
    bn = torch.nn.BatchNorm2d(8).to(cuda).train()
    y = torch.zeros(1, 1, 1, 8, device=cuda)
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        nets.BNOp(bn).forward(y, nets._bn_scratch(8, cuda), True, torch.empty_like(y), ACT_RELU)


def test_fused_loss_against_reference_golden(cuda, lib, golden):
    """every AdversarialLoss branch value + gradient vs fixtures produced by the reference's src/loss.py."""
    from stcgan_b200 import AdversarialLoss, DataLoss
    cr = torch.tensor(golden["adv"]["C_real"]).to(cuda); cf = torch.tensor(golden["adv"]["C_fake"]).to(cuda)
    for ls, rel, avg, d in itertools.product((0, 1), repeat=4):
        key = f"ls{ls}_rel{rel}_avg{avg}_D{d}"
        a, b = cr.clone().requires_grad_(True), cf.clone().requires_grad_(True)
        v = AdversarialLoss(ls=bool(ls), rel=bool(rel), avg=bool(avg)).to(cuda)(a, b, D_loss=bool(d))
        v.backward()
        assert abs(v.item() - float(golden["adv"][key])) < 1e-6 * max(1, abs(float(golden["adv"][key]))), key
        for grad, name in ((a.grad, "dreal"), (b.grad, "dfake")):
            ref = torch.tensor(golden["adv"][f"{key}/{name}"])
            got = torch.zeros_like(ref) if grad is None else grad.cpu()
            assert (got - ref).abs().max().item() < 1e-7, (key, name)
    p = torch.randn(2, 3, 17, 19, device=cuda, requires_grad=True); t = torch.randn(2, 3, 17, 19, device=cuda)
    t.view(-1)[::7] = p.detach().view(-1)[::7]      # exact ties: sign(0) = 0 like F.l1_loss
    v = DataLoss()(p, t); v.backward()
    pc = p.detach().cpu().double().requires_grad_(True)
    rv = F.l1_loss(pc, t.cpu().double()); rv.backward()
    assert abs(v.item() - rv.item()) < 1e-6 and (p.grad.cpu().double() - pc.grad).abs().max().item() < 1e-9


def test_fused_adam_matches_torch_adam(cuda, lib):
    """torch.optim.Adam semantics (betas (0.5, 0.999), eps 1e-8 -- src/cgan.py:85-90) over several steps,
    with gradients in parameter layout and in packed [16][d0][d1] layout."""
    from stcgan_b200 import FusedAdam, ops
    g = torch.Generator().manual_seed(0)
    shapes = [(8, 12, 4, 4), (5,), (16, 4, 4, 4), (3,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    ref = [p.clone().double().requires_grad_(True) for p in ps]
    mine = [torch.nn.Parameter(p.clone().to(cuda)) for p in ps]
    o_ref = torch.optim.Adam(ref, lr=5e-4, betas=(0.5, 0.999))
    o_mine = FusedAdam(mine, lr=5e-4, betas=(0.5, 0.999))
    packed = {0: torch.zeros(16 * 8 * 12, device=cuda), 2: torch.zeros(16 * 16 * 4, device=cuda)}
    o_mine.set_packed_grads({id(mine[i]): (packed[i], shapes[i][0], shapes[i][1]) for i in packed})
    for step in range(4):
        for i, (r, m) in enumerate(zip(ref, mine)):
            gr = torch.randn(r.shape, generator=g)
            r.grad = gr.double()
            if i in packed:
                packed[i].copy_(gr.permute(2, 3, 0, 1).reshape(-1))
            else:
                m.grad = gr.to(cuda)
        if step == 2:
            for grp in o_ref.param_groups + o_mine.param_groups:
                grp["lr"] = 2e-4                              # ExponentialLR-style change between steps
        o_ref.step(); o_mine.step()
        for r, m in zip(ref, mine):
            assert rel_err(m, r) < 2e-6, step
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 4


def test_float2uint_bit_exact(cuda, lib, golden):
    """integer contract: (clip(a,0,1)*255).astype(uint8) bit-exact against the reference's utils.float2uint."""
    from stcgan_b200 import ops
    v = torch.tensor(golden["f2u"]["inputs"]).to(cuda)
    assert np.array_equal(ops.float2uint(v).cpu().numpy(), golden["f2u"]["outputs"])
    pre = torch.tensor(golden["f2u"]["pre"]).reshape(2, 2, 32, 32).to(cuda)      # tanh-range values
    exp = (np.clip(pre.cpu().numpy() * 0.5 + 0.5, 0, 1) * 255).astype(np.uint8).transpose(0, 2, 3, 1)
    assert np.array_equal(ops.float2uint_hwc(pre).cpu().numpy(), exp)


def test_uint8_input_transform_is_bit_exact(cuda, lib):
    """GPU-side dataset transform == numpy float32 arithmetic of src/utils.py:60-62 + src/dataset.py:152, all 256 levels."""
    from stcgan_b200 import ops
    rs = np.random.RandomState(1)
    img = rs.randint(0, 256, size=(2, 9, 13, 3), dtype=np.uint8)
    img[0, 0, :, 0] = np.arange(13) * 21 % 256
    img.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    ref = ((img.astype(np.float32) / 255).transpose(0, 3, 1, 2) - 0.5) * 2
    assert ref.dtype == np.float32
    got = ops.u8_to_nchw(torch.from_numpy(img).to(cuda)).cpu().numpy()
    assert np.array_equal(got, ref)


def test_layout_kernels(cuda, lib):
    from stcgan_b200 import ops
    a = torch.randn(2, 3, 9, 11, device=cuda); b = torch.randn(2, 1, 9, 11, device=cuda); c = torch.randn(2, 3, 9, 11, device=cuda)
    for dt in (torch.float32, torch.bfloat16):
        p = ops.pack_input([a, b, c], 8, dt)
        ref = torch.cat([a, b, c, torch.zeros(2, 1, 9, 11, device=cuda)], 1).permute(0, 2, 3, 1).to(dt)
        assert torch.equal(p, ref)
        assert torch.equal(ops.nhwc_to_nchw(p), ref.float().permute(0, 3, 1, 2))
        assert torch.equal(ops.nchw_to_nhwc(a, dt), a.permute(0, 2, 3, 1).to(dt))
        g = torch.zeros(2, 3, 9, 11, device=cuda)
        ops.unpack_input_grad(p, 4, 3, g, False); ops.unpack_input_grad(p, 4, 3, g, True)
        assert torch.equal(g, 2 * c.to(dt).float())


THIN_CASES = [
    # kind, cin, cout, n, h, w   (the thin first / last layers of G and D, on the tcgen05 thin paths in bf16 mode)
    ("conv2", 3, 64, 2, 32, 32),      # G1 e1
    ("conv2", 4, 64, 3, 16, 24),      # G2 e1 / D1 c1
    ("conv2", 7, 64, 2, 20, 12),      # D2 c1 (+bias)
    ("convT", 128, 1, 2, 16, 16),     # G1 d1 (+bias, tanh)
    ("convT", 128, 3, 3, 8, 12),      # G2 d1
    ("conv1", 512, 1, 2, 9, 9),       # D c5 (+bias)
    ("conv1", 512, 1, 3, 31, 31),     # D c5 at the 256x256 geometry
    ("conv2", 7, 64, 1, 37, 53),      # odd sizes: the input gradient's last row / column come from one tap only
    ("conv2", 4, 64, 2, 64, 96),      # many overlapping col2im tiles
    ("convT", 64, 3, 1, 40, 50),      # col2im tiles that do not divide the image
    ("convT", 128, 1, 1, 3, 5),       # image smaller than one tile
    ("conv1", 64, 1, 2, 20, 45),      # stride-1 gather over several tiles
    ("conv1", 128, 3, 1, 6, 7),       # three output channels through the stride-1 gather
]


@pytest.mark.parametrize("case", THIN_CASES, ids=lambda c: "-".join(map(str, c)))
def test_thin_layers_on_tensor_cores(cuda, lib, case):
    """thin-K (5-D im2col TMA view of a zero-bordered 8-channel tensor), thin-N (pixel GEMM with the taps in the N
    dimension + in-CTA col2im, stcgan_thin_col2im) and thin wgrad kernels against the torch ops they replace, bf16 inputs,
    identical data."""
    from stcgan_b200 import ops
    from stcgan_b200._lib import ACT_NONE, ACT_TANH
    kind, cin, cout, n, h, w_ = case
    mode, dt = "bf16", torch.bfloat16
    op, w, b = _convop(kind, cin, cout, mode, cuda, bias=True)
    assert op.thin == {"conv2": "cin", "convT": "coutT", "conv1": "cout1"}[kind]
    g = torch.Generator().manual_seed(2)
    x = _round(torch.randn(n, cin, h, w_, generator=g), mode)
    xd = x.double().requires_grad_(True); wd = _round(w, mode).double().requires_grad_(True)
    pre = _ref_forward(kind, xd, wd, b.double())
    oh, ow = pre.shape[2:]
    go = _round(torch.randn(pre.shape, generator=g), mode)
    if kind == "conv2":
        xb = ops.pack_input([x.to(cuda)], 8, dt, border=1)
        out = op.forward(None, oh, ow, x_bordered=xb)
        assert rel_err(_nchw(out), pre) < TOL[mode], "thin-K forward"
        pre.backward(go.double())
        d8 = torch.empty(n, h, w_, 8, dtype=dt, device=cuda)
        op.dgrad(_nhwc(go, dt, cuda), h, w_, out8=d8)
        assert rel_err(_nchw(d8[..., :cin]), xd.grad) < TOL[mode], "thin-N dgrad"
        op.wgrad(None, _nhwc(go, dt, cuda), x_bordered=xb)
        assert rel_err(op.gb, go.double().sum(dim=(0, 2, 3))) < 1e-5
    else:
        ref = torch.tanh(pre) if kind == "convT" else pre
        out = torch.empty(n, cout, oh, ow, device=cuda)
        op.forward(_nhwc(x, dt, cuda), oh, ow, out_nchw=out, act=ACT_TANH if kind == "convT" else ACT_NONE)
        assert rel_err(out, ref) < 1e-5, "thin-N forward (fp32 NCHW output: no output rounding)"
        pre.backward(go.double())
        gb = ops.out_act_bwd(ACT_NONE, torch.zeros_like(go).to(cuda), go.to(cuda), dt, cpad=8, border=1 if kind == "convT" else 2)
        gx = op.dgrad(None, h, w_, g_bordered=gb)
        assert rel_err(_nchw(gx), xd.grad) < TOL[mode], "thin-K dgrad"
        op.wgrad(_nhwc(x, dt, cuda), None, g_bordered=gb)
        assert rel_err(op.gb, go.double().sum(dim=(0, 2, 3))) < 1e-5
    gw = ops.unpack_grad(op.g, op.d0, op.d1)
    torch.cuda.synchronize()
    assert rel_err(gw, wd.grad) < 2e-5, "thin wgrad"


@pytest.mark.parametrize("env", [{"STCGAN_TC_PERSISTENT": "1"}, {"STCGAN_TC_PAIR": "1"}, {"STCGAN_TC_BN256": "1"},
                                 {"STCGAN_TC_PERSISTENT": "0"}], ids=lambda e: "-".join(f"{k[10:]}={v}" for k, v in e.items()))
def test_optional_tensor_core_kernels_stay_parity_green(cuda, lib, env):
    """The persistent (TMEM double-buffered), CTA-pair (tcgen05 cta_group::2) and 128x256-tile variants of the tap-GEMM are
    selected by environment variables read once per process, so each runs in a child process: D2's forward + backward
    (c2..c4 exercise every variant: Nout 128 / 256 / 512, stride 2 and 1) against the float64 oracle."""
    import os, subprocess, sys, textwrap
    from conftest import ROOT
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {os.path.join(ROOT, 'shadow-removal-istd_b200')!r}); sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
        import torch, stcgan_b200 as S, stcgan_oracle as O
        st = O.build_all_states()["D2"]
        d = S.NLayerDiscriminator(7); d.load_state_dict(st); d.cuda().train()
        x, m, y = O.make_istd_batch(4, 256, 256)
        inp = torch.cat((x, m, y), 1)
        xi = inp.cuda().requires_grad_(True)
        out = d(xi); out.sum().backward(); torch.cuda.synchronize()
        sd = {{k: (v.double() if v.is_floating_point() else v.clone()) for k, v in st.items()}}
        for k in O.trainable_keys(sd): sd[k].requires_grad_(True)
        xr = inp.double().requires_grad_(True)
        ref = O.discriminator_forward(sd, xr, training=True); ref.sum().backward()
        rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
        e_out, e_dx = rel(out, ref), rel(xi.grad, xr.grad)
        e_w = max(rel(p.grad, sd[k].grad) for k, p in d.named_parameters())
        print("ERR", e_out, e_dx, e_w)
        assert e_out < 2e-2 and e_dx < 0.5 and e_w < 0.5
    """)
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]


EP_CASES = [
    # kind, cin, cout, n, h, w, crop (rows / columns dropped from the output extent)
    ("conv2", 64, 128, 4, 32, 32, (0, 0)),      # encoder level: BN + LeakyReLU / ReLU, two outputs
    ("conv2", 128, 64, 3, 15, 21, (0, 0)),      # odd input (TMA zero fill), BN = 64 tiles
    ("conv2", 64, 128, 6, 128, 128, (0, 0)),    # two-accumulator tiles (MT = 2)
    ("conv2", 512, 512, 4, 4, 6, (0, 0)),       # split-K path: the finisher applies scale / shift / both activations
    ("convT", 128, 64, 2, 16, 16, (0, 0)),      # decoder level
    ("convT", 1024, 512, 2, 8, 10, (1, 0)),     # decoder level with the odd-size crop (16x20 -> 15x20), 128x256-tile candidate
    ("convT", 1024, 512, 2, 2, 3, (0, 1)),      # split-K + crop (4x6 -> 4x5)
    ("convT", 512, 256, 64, 15, 20, (0, 0)),    # inference-sized launch (multi-wave)
]


@pytest.mark.parametrize("case", EP_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_inference_epilogue_folds_eval_batchnorm(cuda, lib, case):
    """stcgan_tapconv_ep: eval-mode BatchNorm (per-channel scale / shift from the running statistics) + two activations + the
    odd-size crop folded into the convolution's epilogue == nn.Conv2d / nn.ConvTranspose2d -> nn.BatchNorm2d.eval() ->
    LeakyReLU / ReLU -> [:, :, :H, :W] of the reference (src/models/stcgan_g.py:85-90, 107-111, 131)."""
    from stcgan_b200 import ops
    from stcgan_b200._lib import ACT_LEAKY, ACT_RELU, GEOM_PARITY, GEOM_WIN_S2
    kind, cin, cout, n, h, w_, (dh, dw) = case
    op, w, _ = _convop(kind, cin, cout, "bf16", cuda)
    g = torch.Generator().manual_seed(7)
    x = _round(torch.randn(n, cin, h, w_, generator=g), "bf16")
    ph, pw = (h + h % 2, w_ + w_ % 2) if kind == "conv2" else (h, w_)
    oh, ow = op.out_size(ph, pw)
    hc, wc = oh - dh, ow - dw
    bn = torch.nn.BatchNorm2d(cout).double().eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(cout, generator=g) + 0.5); bn.bias.copy_(torch.randn(cout, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(cout, generator=g) * 0.2); bn.running_var.copy_(torch.rand(cout, generator=g) + 0.5)
    f32 = lambda t: t.detach().float().to(cuda)
    ss = torch.empty((2, cout), device=cuda); mi = torch.empty((2, cout), device=cuda)
    ops.bn_finalize(None, 0, f32(bn.weight), f32(bn.bias), f32(bn.running_mean), f32(bn.running_var), 0.0, 1e-5, False, mi, ss)
    wide = torch.zeros((n, hc, wc, 2 * cout), dtype=torch.bfloat16, device=cuda)      # out2 = a channel-slice view
    geom = GEOM_PARITY if kind == "convT" else GEOM_WIN_S2
    wp = op.p2 if kind == "convT" else op.p1
    a = ops.tapconv_ep(geom, _nhwc(x, torch.bfloat16, cuda), wp, cout, oh, ow, scale_shift=ss, act=ACT_LEAKY,
                       act2=ACT_RELU, out2=wide[..., cout:], crop=(hc, wc))
    torch.cuda.synchronize()
    xp = F.pad(x.double(), (0, pw - w_, 0, ph - h)) if kind == "conv2" else x.double()
    z = bn(_ref_forward(kind, xp, _round(w, "bf16").double(), None))[:, :, :hc, :wc]
    assert tuple(a.shape) == (n, hc, wc, cout)
    assert rel_err(_nchw(a), F.leaky_relu(z, 0.2)) < 6e-3
    assert rel_err(_nchw(wide[..., cout:]), F.relu(z)) < 6e-3
    assert float(wide[..., :cout].abs().max()) == 0.0                                  # the other half of the buffer is untouched
    # single output, plain shift (a bias) instead of BatchNorm
    b = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    r = ops.tapconv_ep(geom, _nhwc(x, torch.bfloat16, cuda), wp, cout, oh, ow, shift=b, act=ACT_RELU, crop=(hc, wc))
    torch.cuda.synchronize()
    ref = F.relu(_ref_forward(kind, xp, _round(w, "bf16").double(), b.double().cpu()))[:, :, :hc, :wc]
    assert rel_err(_nchw(r), ref) < 6e-3


@pytest.mark.parametrize("cin,n,h,w_", [(3, 2, 64, 64), (4, 3, 30, 40)])
def test_thin_first_layer_with_two_activations(cuda, lib, cin, n, h, w_):
    """stcgan_thinconv2: the U-Net's first Conv2d with both consumers' activations (LeakyReLU for the next down conv, ReLU for
    the skip concatenation, src/models/stcgan_g.py:87,89,125) from one accumulator == the two-pass form."""
    from stcgan_b200 import ops
    from stcgan_b200._lib import ACT_LEAKY, ACT_RELU
    op, w, _ = _convop("conv2", cin, 64, "bf16", cuda)
    assert op.thin == "cin"
    g = torch.Generator().manual_seed(3)
    x = _round(torch.randn(n, cin, h, w_, generator=g), "bf16")
    xb = ops.pack_input([x.to(cuda).contiguous()], 8, torch.bfloat16, border=1)
    oh, ow = op.out_size(h, w_)
    a = torch.empty((n, oh, ow, 64), dtype=torch.bfloat16, device=cuda)
    cat = torch.zeros((n, oh, ow, 128), dtype=torch.bfloat16, device=cuda)
    ops.thinconv2(xb, 2, op.wthin, 64, oh, ow, act=ACT_LEAKY, out=a, act2=ACT_RELU, out2=cat[..., :64])
    y = ops.thinconv(xb, 2, op.wthin, 64, oh, ow)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), _round(w, "bf16").double(), None, 2, 1)
    assert rel_err(_nchw(a), F.leaky_relu(ref, 0.2)) < 6e-3 and rel_err(_nchw(cat[..., :64]), F.relu(ref)) < 6e-3
    assert torch.equal(cat[..., :64], F.relu(y))             # (ReLU commutes with the bf16 rounding; 0.2 * v does not)
    assert float(cat[..., 64:].abs().max()) == 0.0


@pytest.mark.parametrize("case,kw", [("full", dict(scale=0.05, angle=15, flip_prob=0.5)),
                                     ("flipcrop", dict(scale=None, angle=None, flip_prob=0.5))])
def test_gpu_augmentation_matches_reference_golden(cuda, lib, case, kw):
    """stcgan_augment_u8 (RandomScale -> RandomRotate -> RandomHorizontalFlip -> RandomCrop + the dataset transform, SURVEY
    8f-2) against vectors produced by the reference's transform classes on OpenCV (tests/golden/make_golden_augment.py) and
    against the CPU restatement: BIT-EXACT (cv::warpAffine works in fixed-point source positions and a fixed float32 expression)."""
    import os
    import augment_oracle as AO
    from conftest import ROOT
    from stcgan_b200 import augment as A
    vec = np.load(os.path.join(ROOT, "tests", "golden", "augment_vectors.npz"))
    h, w, crop, n = (int(v) for v in vec["meta"])
    np.random.seed(42)
    params = A.sample_params(np.random, n, h, w, crop=crop, **kw)
    for key in ("img", "matte"):
        u8 = torch.from_numpy(vec[f"{case}/{key}_u8"]).to(cuda).contiguous()
        got = A.augment_u8(u8, params, crop).cpu().numpy()
        want = vec[f"{case}/{key}_out"]
        mine = np.stack([AO.augment(vec[f"{case}/{key}_u8"][i], params[i], crop) for i in range(n)])
        assert got.shape == want.shape
        assert np.array_equal(mine, want)       # the CPU restatement is pinned to the reference + OpenCV
        assert np.array_equal(got, want), float(np.abs(got - want).max())      # and the kernel reproduces it bit for bit


@pytest.mark.parametrize("cout,n,h,w_", [(1, 2, 16, 20), (3, 3, 15, 9)])
def test_last_layer_epilogue_quantises_to_uint8(cuda, lib, cout, n, h, w_):
    """stcgan_thin_convt_u8: ConvTranspose2d(128 -> 1|3) + bias + Tanh with utils.float2uint (src/utils.py:65-67, as applied by
    CGAN.infer, src/cgan.py:441-446) inside the epilogue == quantising the same launch's float output with numpy, bit for bit
    -- with and without the float output."""
    import stcgan_oracle as O
    from stcgan_b200 import ops
    from stcgan_b200._lib import ACT_TANH
    op, w, b = _convop("convT", 128, cout, "bf16", cuda, bias=True)
    assert op.thin == "coutT"
    g = torch.Generator().manual_seed(2)
    x = _nhwc(_round(torch.randn(n, 128, h, w_, generator=g) * 0.5, "bf16"), torch.bfloat16, cuda)
    out = torch.empty((n, cout, 2 * h, 2 * w_), dtype=torch.float32, device=cuda)
    u8 = ops.thin_convT_u8(x, op.wtn, op.cpad, cout, 2 * h, 2 * w_, bias=op.bias.detach(), act=ACT_TANH, out_nchw=out)
    u8_only = ops.thin_convT_u8(x, op.wtn, op.cpad, cout, 2 * h, 2 * w_, bias=op.bias.detach(), act=ACT_TANH)
    plain = torch.empty_like(out)
    op.forward(x, 2 * h, 2 * w_, out_nchw=plain, act=ACT_TANH)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)
    want = O.float2uint(out.cpu().numpy().transpose(0, 2, 3, 1) * 0.5 + 0.5)
    assert np.array_equal(u8.cpu().numpy(), want) and np.array_equal(u8_only.cpu().numpy(), want)
    assert 0 < int(u8.max()) and int(u8.min()) < 255 and u8.float().std() > 1          # (a non-degenerate image)
