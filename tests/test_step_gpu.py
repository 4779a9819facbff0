"""Full train-step parity (BASELINE configs[0] and configs[1] restated at test size) and inference parity.

The oracle's train step (oracle/stcgan_oracle.py::OracleTrainer.train_step) is the restatement of
src/cgan.py:274-351 that `oracle/pin_against_reference.py` pins to the reference at 1e-13 in float64.
"""
import numpy as np
import pytest
import torch

import stcgan_oracle as O
from conftest import rel_err

pytestmark = pytest.mark.gpu

OUT_TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _build(mode, dev, states):
    import stcgan_b200 as S
    nets = dict(G1=S.UnetGenerator(3, 1, precision=mode), G2=S.UnetGenerator(4, 3, precision=mode),
                D1=S.NLayerDiscriminator(4, precision=mode), D2=S.NLayerDiscriminator(7, precision=mode))
    for n, mod in nets.items():
        mod.load_state_dict(states[n]); mod.to(dev).train()
    return nets


@pytest.fixture(scope="module")
def states():
    return O.build_all_states()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_engine_train_step_vs_oracle(cuda, lib, states, mode):
    import stcgan_b200 as S
    nets = _build(mode, cuda, states)
    eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"], S.TrainConfig())
    x, m, y = O.make_istd_batch(2, 256, 256, seed=42)
    ref = O.OracleTrainer(states, O.HyperParams(), dtype=torch.float64)
    r = ref.train_step(x.double(), m.double(), y.double(), keep_grads=True)
    if mode == "bf16":      # rounding-noise yardstick: the same oracle with bf16-rounded convolution operands / gradients
        with O.emulate_conv_precision(torch.bfloat16):
            emu = O.OracleTrainer(states, O.HyperParams(), dtype=torch.float64).train_step(
                x.double(), m.double(), y.double(), keep_grads=True)
    eng.train_step(x.to(cuda), m.to(cuda), y.to(cuda))
    torch.cuda.synchronize()
    L = eng.loss_dict()
    tol = OUT_TOL[mode]
    assert rel_err(eng.last["m_pred"], r["m_pred"]) < tol and rel_err(eng.last["y_pred"], r["y_pred"]) < tol
    for k in ("D1_loss", "D2_loss", "D_loss", "data1_loss", "data2_loss"):
        assert abs(L[k] - float(r[k])) <= tol * abs(float(r[k])), (k, L[k], float(r[k]))
    # G-phase adversarial terms are computed with the Adam-updated D (lr*sign(g) amplifies gradient noise)
    for k in ("G1_loss", "G2_loss", "G_loss"):
        assert abs(L[k] - float(r[k])) <= max(tol, 2e-2) * abs(float(r[k])), (k, L[k], float(r[k]))
    # D gradients are taken before any optimiser step: compare the packed gradients
    from stcgan_b200 import ops
    for n in ("D1", "D2"):
        rt = eng.rt[n]
        for i, (p, gref) in enumerate(zip(nets[n].parameters(), r["grads_D"][n])):
            v, d0, d1 = rt.param_grad_views[id(p)]
            got = ops.unpack_grad(v, d0, d1) if d0 else v.view(p.shape)
            e = rel_err(got, gref)
            bound = 5e-3 if mode == "fp32" else 1.5 * max(rel_err(emu["grads_D"][n][i], gref), 2e-2)
            assert e < bound, (n, tuple(p.shape), e, bound)
    # BN running statistics: D saw 4 training forwards, G one (cgan.py:281-289, 321-324)
    for n in nets:
        for k, v in nets[n].state_dict().items():
            if "num_batches" in k:
                assert int(v) == int(ref.sd[n][k]) == (4 if n[0] == "D" else 1)
            if "running" in k:
                assert rel_err(v, ref.sd[n][k]) < max(tol, 2e-2), (n, k)
    # the optimiser kernel also refreshed the packed bf16 weight copies (no separate pack pass): check them
    if mode == "bf16":
        for n in nets:
            for conv in eng.rt[n].convs:
                w = conv.weight.detach()
                assert torch.equal(conv.p1, w.permute(2, 3, 0, 1).reshape(16, conv.d0, conv.d1).to(torch.bfloat16)), n
                assert torch.equal(conv.p2, w.permute(2, 3, 1, 0).reshape(16, conv.d1, conv.d0).to(torch.bfloat16)), n
    # post-Adam parameters: |delta| <= lr per element in the first step; compare the update direction statistically
    for n, lr in (("G1", 5e-4), ("G2", 5e-4), ("D1", 1e-4), ("D2", 1e-4)):
        agree, total = 0, 0
        for (k, p), q in zip(nets[n].named_parameters(), ref.params[n]):
            dm = (p.detach().cpu().double() - states[n][k].double()); dr = (q.detach() - states[n][k].double())
            assert dm.abs().max().item() <= lr * 1.001
            agree += int((torch.sign(dm) == torch.sign(dr)).sum()); total += dm.numel()
        assert agree / total > (0.99 if mode == "fp32" else 0.80), (n, agree / total)


def test_engine_matches_autograd_modules(cuda, lib, states):
    """the drop-in nn.Modules driven by torch autograd + the loss modules + FusedAdam, in the order of
    src/cgan.py:274-351, give the same step as the hand-scheduled engine (fp32 mode, same kernels)."""
    import stcgan_b200 as S
    x, m, y = (t.to(cuda) for t in O.make_istd_batch(2, 256, 256, seed=7))
    cfg = S.TrainConfig()
    a = _build("fp32", cuda, states)
    eng = S.STCGANEngine(a["G1"], a["G2"], a["D1"], a["D2"], cfg)
    eng.train_step(x, m, y)
    La = eng.loss_dict()
    b = _build("fp32", cuda, states)
    G1, G2, D1, D2 = b["G1"], b["G2"], b["D1"], b["D2"]
    optG = S.FusedAdam(list(G1.parameters()) + list(G2.parameters()), lr=cfg.lr_G, betas=(cfg.beta1, cfg.beta2))
    optD = S.FusedAdam(list(D1.parameters()) + list(D2.parameters()), lr=cfg.lr_D, betas=(cfg.beta1, cfg.beta2))
    adv, dl = S.AdversarialLoss().to(cuda), S.DataLoss()
    optD.zero_grad(); optG.zero_grad()
    C1r = D1(torch.cat((x, m), 1)); mp = G1(x); C1f = D1(torch.cat((x, mp.detach()), 1))
    C2r = D2(torch.cat((x, m, y), 1)); yp = G2(torch.cat((x, mp), 1)); C2f = D2(torch.cat((x, mp.detach(), yp.detach()), 1))
    D1l, D2l = adv(C1r, C1f, D_loss=True), adv(C2r, C2f, D_loss=True)
    (cfg.lambda2 * D1l + cfg.lambda3 * D2l).backward(); optD.step()
    optG.zero_grad(); D1.requires_grad_(False); D2.requires_grad_(False)
    C1r = D1(torch.cat((x, m), 1)); C1f = D1(torch.cat((x, mp), 1))
    C2r = D2(torch.cat((x, m, y), 1)); C2f = D2(torch.cat((x, mp, yp), 1))
    G1l, G2l = adv(C1r, C1f, D_loss=False), adv(C2r, C2f, D_loss=False)
    d1, d2 = dl(mp, m), dl(yp, y)
    Gl = d1 + cfg.lambda1 * d2 + cfg.lambda2 * G1l + cfg.lambda3 * G2l
    Gl.backward(); optG.step()
    torch.cuda.synchronize()
    for k, v in (("D1_loss", D1l), ("D2_loss", D2l), ("G1_loss", G1l), ("G2_loss", G2l), ("data1_loss", d1),
                 ("data2_loss", d2), ("G_loss", Gl)):
        # the G-phase adversarial terms are evaluated with the Adam-updated discriminators: +-lr_D flips of noise-level
        # gradients (atomics order) move them by a few 1e-4 between any two runs
        tol = 2e-3 if k in ("G1_loss", "G2_loss", "G_loss") else 1e-4
        assert abs(La[k] - v.item()) <= tol * abs(v.item()) + 1e-7, (k, La[k], v.item())
    for n in a:
        for (k, p), q in zip(a[n].named_parameters(), b[n].parameters()):
            d = (p - q).abs()
            lr = cfg.lr_G if n[0] == "G" else cfg.lr_D
            # Adam's first step is lr*sign(g): noise-level gradients (atomics ordering) may flip by up to 2*lr; a BatchNorm
            # bias gradient is a near-cancelling sum over all pixels, so a few per cent of its few hundred entries may flip
            frac = (d > 0.1 * lr).float().mean().item()
            assert d.max().item() <= 2.02 * lr and frac < (0.02 if d.numel() > 1024 else 0.06), (n, k, frac)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_inference_and_quantisation_vs_oracle(cuda, lib, states, mode, golden):
    """configs[3] geometry (480x640, eval BN) at batch 1 + the uint8 contract of utils.float2uint."""
    import stcgan_b200 as S
    nets = _build(mode, cuda, states)
    nets["G1"].eval(); nets["G2"].eval()
    x = O.make_istd_batch(1, 480, 640, seed=5)[0]
    mp, yp, m8, y8 = S.infer(nets["G1"], nets["G2"], x.to(cuda))
    om, oy, om8, oy8 = O.infer(states["G1"], states["G2"], x)
    tol = OUT_TOL[mode]
    assert rel_err(mp, om) < tol and rel_err(yp, oy) < tol
    # golden samples produced by the reference itself
    s = int(golden["step"]["stride"])
    assert rel_err(mp.cpu().reshape(-1)[::s], torch.tensor(golden["infer"]["m_pred/sample"])) < tol
    assert rel_err(yp.cpu().reshape(-1)[::s], torch.tensor(golden["infer"]["y_pred/sample"])) < tol
    # bit-exact integer contract: quantising OUR floats on the GPU == numpy float2uint of the same floats
    assert np.array_equal(m8.cpu().numpy()[0], O.float2uint(mp.cpu().numpy()[0].transpose(1, 2, 0) * 0.5 + 0.5))
    assert np.array_equal(y8.cpu().numpy()[0], O.float2uint(yp.cpu().numpy()[0].transpose(1, 2, 0) * 0.5 + 0.5))
    # and the images agree with the reference's to within one grey level almost everywhere
    diff = np.abs(y8.cpu().numpy()[0].astype(int) - oy8[0].astype(int))
    assert diff.max() <= (1 if mode == "fp32" else 6) and (diff > 0).mean() < (0.02 if mode == "fp32" else 0.6)


def test_infer_u8_end_to_end_equals_float_path(cuda, lib, states):
    """stcgan_b200.infer_u8 (uint8 HWC images in, uint8 HWC images out, pinned host buffers) == the float path fed with the
    dataset transform of the same images (src/dataset.py:100-110,152)."""
    import stcgan_b200 as S
    nets = _build("bf16", cuda, states)
    nets["G1"].eval(); nets["G2"].eval()
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (2, 96, 128, 3), generator=g, dtype=torch.uint8)
    # the reference's host-side transform in numpy float32 (src/utils.py:60-62, src/dataset.py:152)
    x = torch.from_numpy(np.ascontiguousarray(((img.numpy().astype(np.float32) / 255).transpose(0, 3, 1, 2) - 0.5) * 2))
    _, _, m8_ref, y8_ref = S.infer(nets["G1"], nets["G2"], x.to(cuda))
    out_m = torch.empty((2, 96, 128, 1), dtype=torch.uint8).pin_memory()
    out_y = torch.empty((2, 96, 128, 3), dtype=torch.uint8).pin_memory()
    m8, y8 = S.infer_u8(nets["G1"], nets["G2"], img.pin_memory(), out_m, out_y)
    torch.cuda.synchronize()
    # (the bottleneck layers reduce their split-K partial sums with fp32 atomics in arbitrary order, so two runs of the same
    # network agree only up to a bf16 ulp, i.e. an occasional grey level)
    for a, b in ((m8, m8_ref), (y8, y8_ref)):
        d = (a.int() - b.int()).abs()
        assert int(d.max()) <= 2 and float((d > 0).float().mean()) < 0.05
    assert torch.equal(out_m, m8.cpu()) and torch.equal(out_y, y8.cpu())


def test_replay_async_pipelines_inputs_and_loss_reads(cuda, lib, states):
    """STCGANEngine.replay_async: host batches in through the double-buffered inbox, losses out one step later; the values
    are those of the plain replay path (same captured graph, same inputs)."""
    import stcgan_b200 as S
    nets = _build("bf16", cuda, states)
    eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"])
    batches = [tuple(t.contiguous().pin_memory() for t in O.make_istd_batch(2, 256, 256, seed=s)) for s in (1, 2, 3)]
    eng.capture(*(t.to(cuda) for t in batches[0]), warmup=1)
    assert eng.replay_async(*batches[0]) is None
    l0 = eng.replay_async(*batches[1])                   # losses of the step on batches[0]
    l1 = eng.replay_async(*batches[2])
    l2 = eng.flush()
    torch.cuda.synchronize()
    for l in (l0, l1, l2):
        assert l is not None and not l.is_cuda and torch.isfinite(l).all()
    assert torch.equal(l2, eng.losses.cpu())             # the last step's losses are what the device holds
    assert not torch.equal(l0, l1) and not torch.equal(l1, l2)
    assert float(eng.optim_G.state_dict()["state"][0]["step"]) == 4     # 1 warm-up + 3 replays (capturing is not a step)


def test_inference_pipeline_equals_infer_u8(cuda, lib, states):
    """stcgan_b200.InferencePipeline (transfers overlapped with compute, results one step later) returns exactly what the
    serial infer_u8 returns, batch by batch."""
    import stcgan_b200 as S
    nets = _build("bf16", cuda, states)
    nets["G1"].eval(); nets["G2"].eval()
    g = torch.Generator().manual_seed(4)
    batches = [torch.randint(0, 256, (2, 96, 128, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(4)]
    want = []
    for b in batches:
        m8, y8 = S.infer_u8(nets["G1"], nets["G2"], b)
        want.append((m8.cpu(), y8.cpu()))
    pipe = S.InferencePipeline(nets["G1"], nets["G2"])
    got = []
    for b in batches:
        r = pipe.submit(b)
        if r is not None:
            got.append((r[0].clone(), r[1].clone()))
    r = pipe.flush()
    got.append((r[0].clone(), r[1].clone()))
    assert len(got) == len(want)
    for (gm, gy), (wm, wy) in zip(got, want):       # (split-K atomics order: an occasional grey level between two runs)
        for a, b in ((gm, wm), (gy, wy)):
            d = (a.int() - b.int()).abs()
            assert int(d.max()) <= 2 and float((d > 0).float().mean()) < 0.05
