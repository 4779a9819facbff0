"""CPU-side tests (run with -m "not gpu"): the oracle against the golden fixtures produced by the reference, the C-ABI
library's exports, and the host logic of the package.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import stcgan_oracle as O
from conftest import ROOT, rel_err


# ---------------------------------------------------------------------------------------------------------------
# oracle vs the reference's own outputs (tests/golden, written by tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def states():
    return O.build_all_states()


def test_oracle_weights_match_reference_fingerprint(states, golden):
    for key in golden["weights"].files:
        n, k = key.split("/", 1)
        v = states[n][k].double()
        assert np.allclose([v.sum().item(), v.abs().sum().item()], golden["weights"][key], rtol=0, atol=0), key
    assert sum(v.numel() for k, v in states["G1"].items() if k in O.trainable_keys(states["G1"])) == 54_409_857
    assert sum(v.numel() for k, v in states["D2"].items() if k in O.trainable_keys(states["D2"])) == 2_769_729


def test_oracle_train_step_matches_reference_golden(states, golden):
    """float64 oracle step vs float64 reference step: 1e-9; this is the rigorous pin that travels to the GPU box."""
    g = golden["step"]
    s = int(g["stride"])
    tr = O.OracleTrainer(states, O.HyperParams(), dtype=torch.float64)
    x, m, y = (t.double() for t in O.make_istd_batch(2, 256, 256, seed=42))
    r = tr.train_step(x, m, y, keep_grads=True)
    samp = lambda t: t.detach().reshape(-1)[::s] if t.numel() > 4096 else t.detach().reshape(-1)
    for k in ("m_pred", "y_pred", "C1_fake_Gphase", "C2_fake_Gphase"):
        assert rel_err(samp(r[k]), torch.tensor(g[f"f64/{k}/sample"])) < 1e-9, k
        assert abs(r[k].norm().item() - float(g[f"f64/{k}/norm"])) < 1e-9 * float(g[f"f64/{k}/norm"])
    for k in ("D1_loss", "D2_loss", "D_loss", "G1_loss", "G2_loss", "data1_loss", "data2_loss", "G_loss"):
        assert abs(float(r[k]) - float(g[f"f64/{k}"])) < 1e-10 * max(1.0, abs(float(g[f"f64/{k}"]))), k
    for grp in ("grads_D", "grads_G"):
        for n in r[grp]:
            for k, gr in zip(O.trainable_keys(tr.sd[n]), r[grp][n]):
                assert rel_err(samp(gr), torch.tensor(g[f"f64/grad/{n}/{k}/sample"])) < 1e-8, (n, k)
    for n in tr.sd:
        for k, v in tr.sd[n].items():
            if "running" in k:
                assert rel_err(v, torch.tensor(g[f"f64/post/{n}/{k}"])) < 1e-10, (n, k)
            elif "num_batches" in k:
                assert int(v) == int(g[f"f64/post/{n}/{k}"])
            else:
                assert rel_err(samp(v), torch.tensor(g[f"f64/post/{n}/{k}/sample"])) < 1e-9, (n, k)


def test_oracle_losses_match_reference_golden(golden):
    a = golden["adv"]
    cr, cf = torch.tensor(a["C_real"]), torch.tensor(a["C_fake"])
    for ls in (0, 1):
        for rel in (0, 1):
            for avg in (0, 1):
                for d in (0, 1):
                    key = f"ls{ls}_rel{rel}_avg{avg}_D{d}"
                    x, y = cr.clone().requires_grad_(True), cf.clone().requires_grad_(True)
                    v = O.adversarial_loss(x, y, bool(d), bool(ls), bool(rel), bool(avg))
                    v.backward()
                    assert v.item() == float(a[key]), key
                    assert torch.equal(x.grad if x.grad is not None else torch.zeros_like(cr), torch.tensor(a[key + "/dreal"]))
                    assert torch.equal(y.grad if y.grad is not None else torch.zeros_like(cf), torch.tensor(a[key + "/dfake"]))


def test_oracle_float2uint_and_inference_golden(states, golden):
    f = golden["f2u"]
    assert np.array_equal(O.float2uint(f["inputs"]), f["outputs"])
    x = O.make_istd_batch(1, 480, 640, seed=5)[0]
    mp, yp, m8, y8 = O.infer(states["G1"], states["G2"], x)
    s = int(golden["step"]["stride"])
    gi = golden["infer"]
    assert rel_err(mp.reshape(-1)[::s], torch.tensor(gi["m_pred/sample"])) < 2e-5
    assert rel_err(yp.reshape(-1)[::s], torch.tensor(gi["y_pred/sample"])) < 2e-5
    assert (m8.reshape(-1)[::s] != gi["m_u8/sample"]).mean() < 1e-2 and (y8.reshape(-1)[::s] != gi["y_u8/sample"]).mean() < 1e-2


def test_oracle_flop_model_matches_survey():
    assert abs(O.train_step_flops(256, 256) / 1e9 - 186.23) < 0.01
    assert abs(O.train_step_flops(512, 512) / 1e9 - 754.44) < 0.01
    assert abs(O.inference_flops(480, 640) / 1e9 - 113.29) < 0.01


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_oracle_state_and_forward_match_imported_reference(states):
    """in the build container the oracle is also checked against the live reference modules (small, fast subset of
    oracle/pin_against_reference.py)."""
    import pin_against_reference as P
    networks, loss, utils = P.load_reference()
    nets = P.build_reference_nets(networks)
    for n in nets:
        rsd = nets[n].state_dict()
        assert list(rsd.keys()) == list(states[n].keys())
        assert all(torch.equal(rsd[k], states[n][k]) for k in rsd)
    x, m, _ = O.make_istd_batch(1, 256, 256)
    sd = {k: v.clone() for k, v in states["D1"].items()}
    nets["D1"].train()
    with torch.no_grad():
        assert rel_err(O.discriminator_forward(sd, torch.cat((x, m), 1)), nets["D1"](torch.cat((x, m), 1))) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# the C ABI
# ---------------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "stcgan_b200.h")).read()
    declared = set(re.findall(r"\b(stcgan_[a-z0-9_]+)\s*\(", hdr))
    from stcgan_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.stcgan_abi_version() == 1 and lib.stcgan_arch() == b"sm_100a"
    assert b"invalid argument" in lib.stcgan_error_string(-1) and lib.stcgan_error_string(0) == b"ok"


def test_ctypes_signatures_match_header_prototypes():
    """every prototype of include/stcgan_b200.h has as many parameters as the ctypes binding passes (a drift between the
    header and stcgan_b200/_lib.py would corrupt the stack silently), and every one cites what it replaces nearby"""
    from stcgan_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "stcgan_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\b(stcgan_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", code, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
    assert hdr.count(".py:") >= 30          # reference call sites (file:line) are cited throughout the header


def test_library_argument_validation_without_gpu(lib):
    """bad arguments are rejected on the host before any launch (error convention: negative code, no throw)."""
    assert lib.stcgan_tapconv(99, 0, 0, 1, 1, 4, 4, 8, 8, 1, None, 0, 1, 2, 2, 8, 8, 0, None, 0, None) == -1      # unknown geometry
    assert lib.stcgan_tapconv(0, 7, 0, 1, 1, 4, 4, 8, 8, 1, None, 0, 1, 2, 2, 8, 8, 0, None, 0, None) == -1       # unknown dtype
    assert lib.stcgan_tapconv(0, 0, 1, 1, 1, 4, 4, 64, 64, 1, None, 0, 1, 2, 2, 64, 64, 0, None, 0, None) == -2   # TC backend is bf16-only
    assert lib.stcgan_bn_stats(0, None, 10, 8, 8, None, None) == -1
    assert lib.stcgan_fused_loss(None, 1, None, None) == -1
    assert lib.stcgan_adam_chunk() == 4096


def test_missing_library_fails_loudly(monkeypatch):
    from stcgan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libstcgan_b200.so")
    with pytest.raises(_lib.StcganLibraryError, match="no non-CUDA fallback"):
        _lib.load()


# ---------------------------------------------------------------------------------------------------------------
# host logic of the package
# ---------------------------------------------------------------------------------------------------------------
def test_modules_reproduce_reference_state_dict_and_init(states):
    import stcgan_b200 as S
    torch.manual_seed(O.REFERENCE_SEED)
    nets = dict(G1=S.get_generator("stcgan", in_channels=3, out_channels=1, ngf=64, drop_rate=0.05, no_conv_t=False,
                                   use_selu=False, activation="none"),
                G2=S.get_generator("stcgan", in_channels=4, out_channels=3, ngf=64),
                D1=S.get_discriminator("stcgan", in_channels=4, out_channels=1, ndf=64, use_selu=False, use_sigmoid=False),
                D2=S.get_discriminator("stcgan", in_channels=7, out_channels=3, ndf=64))
    for n, mod in nets.items():
        sd = mod.state_dict()
        assert list(sd.keys()) == list(states[n].keys()), n
        assert all(torch.equal(sd[k], states[n][k]) for k in sd), n           # same RNG consumption as the reference ctor
        assert [k for k, _ in mod.named_parameters()] == O.trainable_keys(states[n])
    assert len(nets["G1"].state_dict()) == 82 and len(nets["D1"].state_dict()) == 22
    torch.manual_seed(7); nets["G1"].apply(S.weights_init)
    ref = O.build_all_states()["G1"]; torch.manual_seed(7); O.apply_weights_init(ref)
    assert all(torch.equal(v, ref[k]) for k, v in nets["G1"].state_dict().items())
    with pytest.raises(KeyError):
        S.get_generator("mnet", in_channels=3, out_channels=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        nets["D1"](torch.zeros(1, 4, 64, 64))


def test_generator_level_sizes_follow_pad_to_even_chain():
    import stcgan_b200 as S
    rt = S.UnetGenerator(3, 1)._build_runtime()
    assert rt.sizes(256, 256) == [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (1, 1)]
    assert rt.sizes(480, 640) == [(480, 640), (240, 320), (120, 160), (60, 80), (30, 40), (15, 20), (8, 10), (4, 5), (2, 3)]


def test_loss_modules_api():
    import stcgan_b200 as S
    a = S.AdversarialLoss(ls=True)
    assert float(a.fake_label) == -1.0 and float(a.real_label) == 1.0 and set(dict(a.named_buffers())) == {"real_label", "fake_label"}
    assert float(S.AdversarialLoss().fake_label) == 0.0
    with pytest.raises(RuntimeError, match="CUDA only"):
        S.DataLoss()(torch.zeros(2, 2), torch.zeros(2, 2))


def test_install_into_reference_shim():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("reference tree not present")
    import subprocess, sys
    code = (f"import sys; sys.path.insert(0, {os.path.join(ROOT, 'shadow-removal-istd_b200')!r}); import stcgan_b200 as S;"
            f"S.install_into_reference({ref!r}); import src.networks as nw;"
            "g = nw.get_generator('stcgan', in_channels=3, out_channels=1, ngf=64, drop_rate=0.0, no_conv_t=False, use_selu=False, activation='none');"
            "d = nw.get_discriminator('stcgan', in_channels=4, out_channels=1, ndf=64, use_selu=False, use_sigmoid=False);"
            "assert type(g) is S.UnetGenerator and type(d) is S.NLayerDiscriminator; print('ok')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_data_parallel_oracle_and_precision_emulation_are_consistent():
    """Round-2 additions to the oracle (test infrastructure of the multi-rank and bf16-bound GPU tests): the single-process
    data-parallel statement with ONE shard is the plain train step, bit for bit; with two shards its gradients are the mean of
    the per-shard gradients (nn.DataParallel semantics, src/cgan.py:78-84); the bf16 emulation only rounds convolution
    operands (it is the identity on values that are already bf16-representable zero tensors, and a small perturbation else)."""
    torch.manual_seed(1)
    st = O.build_all_states(ngf=8, ndf=8)
    sh = [tuple(t.double() for t in O.make_istd_batch(1, 256, 256, seed=3 + i)) for i in range(2)]
    a = O.OracleTrainer(st, dtype=torch.float64)
    ra = a.train_step(*sh[0], keep_grads=True)
    b = O.OracleDataParallel(st, 1, dtype=torch.float64)
    rb, gb = b.train_step([sh[0]])
    assert float(ra["G_loss"]) == float(rb[0]["G_loss"]) and float(ra["D_loss"]) == 0.5 * float(rb[0]["D1_loss"]) + 0.5 * float(rb[0]["D2_loss"])
    for n in a.params:
        for p, q in zip(a.params[n], b.t.params[n]):
            assert torch.equal(p, q)
    # two shards: D-phase gradients == mean of the two single-shard gradient sets
    solo = [O.OracleTrainer(st, dtype=torch.float64).train_step(*s, do_optim=False, keep_grads=True)["grads_D"] for s in sh]
    _, g2 = O.OracleDataParallel(st, 2, dtype=torch.float64).train_step(sh)
    for n in ("D1", "D2"):
        for ga, gb_, gm in zip(solo[0][n], solo[1][n], g2[n]):
            assert torch.allclose((ga + gb_) / 2, gm, rtol=1e-10, atol=1e-14)
    # bf16 emulation: perturbs the step at the bf16 rounding level, nothing more
    with O.emulate_conv_precision(torch.bfloat16):
        re_ = O.OracleTrainer(st, dtype=torch.float64).train_step(*sh[0])
    e = rel_err(re_["y_pred"], ra["y_pred"])
    assert 1e-5 < e < 5e-2, e
