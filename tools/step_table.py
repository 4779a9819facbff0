"""Per-launch table of one eager train step (B=16, 256x256, bf16): every `ops.*` call timed with CUDA events behind a
device-side sleep (so the intervals contain kernel time only), with its shapes, algorithmic FLOPs / bytes and the rate
reached.  Usage: python tools/step_table.py [batch] [H] [W] > table.txt"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stcgan_b200 as S
import stcgan_oracle as O
from stcgan_b200 import ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
W = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda:0")
torch.manual_seed(O.REFERENCE_SEED)
nets = dict(G1=S.UnetGenerator(3, 1), G2=S.UnetGenerator(4, 3), D1=S.NLayerDiscriminator(4), D2=S.NLayerDiscriminator(7))
for n in nets.values():
    n.to(dev).train()
eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"])
for rt in eng.rt.values():
    rt.side_stream = None          # per-launch event intervals need every kernel on the timing stream
eng.lanes.streams = []
eng.hi_stream = None
x, m, y = (t.contiguous().to(dev) for t in O.make_istd_batch(batch, H, W))
for _ in range(3):
    eng.train_step(x, m, y)
torch.cuda.synchronize()


def nb(t):
    return 0 if t is None else t.numel() * t.element_size()


def describe(name, a, k):
    """-> (shape string, flops, bytes)"""
    if name in ("tapconv", "tapconv_thin_n"):
        geom, xx, wp, nout, oh, ow = a[:6]
        n, ih, iw, kk = xx.shape
        taps = 4 if geom == 3 else 16
        fl = 2.0 * n * oh * ow * nout * kk * taps
        by = n * ih * iw * kk * 2 + n * oh * ow * nout * (4 if k.get("out_nchw") is not None else 2) + nb(wp)
        return f"g{geom} x[{n},{ih},{iw},{kk}] -> [{oh},{ow},{nout}]", fl, by
    if name == "tapwgrad":
        geom, s, l, g = a[:4]
        n, sh, sw, d0 = s.shape
        fl = 2.0 * n * sh * sw * d0 * l.shape[3] * 16
        by = s.numel() * 2 + l.numel() * 2 + g.numel() * 4
        return f"g{geom} S[{n},{sh},{sw},{d0}] L[{l.shape[1]},{l.shape[2]},{l.shape[3]}]", fl, by
    if name == "thinconv":
        t, stride, wthin, nout, oh, ow = a[:6]
        fl = 2.0 * t.shape[0] * oh * ow * nout * 128
        by = nb(t) + t.shape[0] * oh * ow * nout * 2
        return f"T[{t.shape[0]},{t.shape[1]},{t.shape[2]},8] s{stride} -> [{oh},{ow},{nout}]", fl, by
    if name == "thinwgrad":
        t, stride, thin_c, f, g = a[:5]
        fl = 2.0 * f.shape[0] * f.shape[1] * f.shape[2] * f.shape[3] * 128
        by = nb(t) + f.numel() * 2
        return f"T[{t.shape[0]},{t.shape[1]},{t.shape[2]},8] s{stride} F[{f.shape[1]},{f.shape[2]},{f.shape[3]}]", fl, by
    if name == "bn_stats":
        yy = a[0]
        return f"y{list(yy.shape)}", 0, yy.numel() * 2
    if name == "bn_act_apply":
        yy, ss, out1, act1 = a[:4]
        out2 = a[4] if len(a) > 4 else k.get("out2")
        by = yy.numel() * 2 + out1.numel() * 2 + (0 if out2 is None else out2.numel() * 2)
        return f"y{list(yy.shape)} ss={'y' if ss is not None else 'n'} out2={'y' if out2 is not None else 'n'}", 0, by
    if name == "bn_fused_apply":
        yy = a[0]; out1 = a[12]; out2 = a[14] if len(a) > 14 else k.get("out2")
        by = yy.numel() * 2 + out1.numel() * 2 + (0 if out2 is None else out2.numel() * 2)
        return f"y{list(yy.shape)} out2={'y' if out2 is not None else 'n'}", 0, by
    if name == "bn_act_bwd":
        yy, ss, mi, gamma, training, g1, act1, g2, act2, acc, dy = a[:11]
        e = dy.numel() * 2
        g2b = 0 if g2 is None else dy.numel() * 2
        by = (2 * e + g2b) * (2 if (ss is not None) else 1) + e          # reduce pass (y, g1[, g2]) + apply pass (y, g1[, g2], dy)
        return f"y{list(yy.shape)} bn={'y' if ss is not None else 'n'} g2={'y' if g2 is not None else 'n'}", 0, by
    if name == "pack_input":
        srcs, cpad, dtype = a[:3]
        n, _, h, w = srcs[0].shape
        b = k.get("border", 0)
        return f"{len(srcs)} src -> [{n},{h + 2 * b},{w + 2 * b},{cpad}]", 0, sum(nb(s) for s in srcs) + n * (h + 2 * b) * (w + 2 * b) * cpad * 2
    if name == "out_act_bwd":
        act, o, do = a[:3]
        return f"out{list(o.shape)}", 0, nb(o) + nb(do) + o.shape[0] * o.shape[2] * o.shape[3] * 16
    if name == "unpack_input_grad":
        g = a[0]
        return f"g{list(g.shape)}", 0, nb(g) // 2
    if name == "colsum":
        return f"g{list(a[0].shape)}", 0, nb(a[0])
    if name == "fused_loss":
        return f"{len(a[0])} terms", 0, sum(nb(t['a']) * 3 for t in a[0])
    if name in ("pack_weight", "pack_weight_thin", "pack_weight_pad16"):
        return f"w{list(a[0].shape)}", 0, nb(a[0]) * 2
    return "", 0, 0


names = ["tapconv", "tapconv_thin_n", "tapwgrad", "thinconv", "thinwgrad", "bn_stats", "bn_finalize", "bn_act_apply", "bn_fused_apply",
         "bn_act_bwd", "pack_input", "out_act_bwd", "unpack_input_grad", "colsum", "fused_loss", "pack_weight",
         "pack_weight_thin", "pack_weight_pad16"]
rec, saved = [], {}


def timed(name, fn):
    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        rec.append((name, describe(name, a, k), e0, e1))
        return out
    return wrapper


for nme in names:
    saved[nme] = getattr(ops, nme)
    setattr(ops, nme, timed(nme, saved[nme]))
# optimiser steps
for oname in ("optim_D", "optim_G"):
    opt = getattr(eng, oname)
    nparam = sum(p.numel() for g in opt.param_groups for p in g["params"])
    opt.step = timed("adam", opt.step)
    opt._table_desc = nparam
_describe = describe


def describe(name, a, k):   # noqa: F811
    if name == "adam":
        return "multi-tensor Adam (+ packed bf16 refresh)", 0, 0
    return _describe(name, a, k)


torch.cuda.synchronize()
torch.cuda._sleep(int(4e9))
ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ea.record()
eng.train_step(x, m, y)
eb.record()
torch.cuda.synchronize()
tot = ea.elapsed_time(eb) * 1e3
print(f"# eager step {tot:.1f} us, {len(rec)} timed op calls, batch {batch} {H}x{W}")
agg = collections.OrderedDict()
tsum = 0.0
for name, (shape, fl, by), e0, e1 in rec:
    us = e0.elapsed_time(e1) * 1e3
    tsum += us
    key = (name, shape)
    c = agg.setdefault(key, [0, 0.0, fl, by])
    c[0] += 1; c[1] += us
print(f"# sum of timed ops {tsum:.1f} us")
print(f"{'op':18s} {'calls':>5s} {'us/call':>9s} {'total us':>9s} {'TF/s':>7s} {'GB/s':>7s}  shape")
for (name, shape), (c, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per = us / c
    print(f"{name:18s} {c:5d} {per:9.1f} {us:9.1f} {fl / per / 1e6 if fl else 0:7.0f} {by / per / 1e3 if by else 0:7.0f}  {shape}")
fam = collections.defaultdict(float)
for (name, shape), (c, us, fl, by) in agg.items():
    fam[name] += us
print("# by op")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f"#   {k:18s} {v:9.1f} us  {100 * v / tsum:5.1f}%")
