"""Timing probe for single tap-GEMM launches (events, 20 repetitions): python tools/conv_probe.py
Environment knobs of the library (STCGAN_TC_DBGMODE etc.) are read per call, so one process can compare them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200"))
import torch
from stcgan_b200 import ops
from stcgan_b200._lib import BACKEND_TC

dev = torch.device("cuda:0")
SHAPES = [  # geom, N, IH, IW, K, OH, OW, Nout, name
    (2, 16, 31, 31, 512, 32, 32, 256, "c4 dgrad (124,4,1) 64 it"),
    (1, 16, 32, 32, 256, 31, 31, 512, "c4 fwd"),
    (3, 16, 64, 64, 128, 128, 128, 64, "c2 dgrad parity N=64 8 it"),
    (0, 16, 128, 128, 64, 64, 64, 128, "e2 fwd 16 it"),
    (0, 16, 64, 64, 128, 32, 32, 256, "e3 fwd 32 it"),
    (0, 16, 32, 32, 256, 16, 16, 512, "e4 fwd lone CTA 64 it"),
    (3, 16, 64, 64, 256, 128, 128, 64, "d2 fwd N=64 16 it"),
    (3, 16, 16, 16, 1024, 32, 32, 256, "d4 fwd parity N=256 64 it"),
    (0, 16, 64, 64, 128, 32, 32, 512, "d3 dgrad N=512 32 it"),
    (3, 16, 32, 32, 256, 64, 64, 128, "c3 dgrad parity N=128 16 it"),
    (3, 16, 32, 32, 512, 64, 64, 128, "d3 fwd parity N=128 32 it"),
]


def run(shape, reps=20):
    geom, n, ih, iw, k, oh, ow, nout, name = shape
    x = torch.randn(n, ih, iw, k, device=dev).to(torch.bfloat16)
    wp = torch.randn(16, nout, k, device=dev).to(torch.bfloat16)
    out = torch.empty(n, oh, ow, nout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.tapconv(geom, x, wp, nout, oh, ow, out=out, backend=BACKEND_TC)
    torch.cuda.synchronize()
    # host-side cost per call (ctypes + 5 tensor-map encodes) exceeds the small kernels: time a captured graph of `reps` launches
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            ops.tapconv(geom, x, wp, nout, oh, ow, out=out, backend=BACKEND_TC)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    taps = 4 if geom == 3 else 16
    fl = 2.0 * n * oh * ow * nout * k * taps
    return us, fl / us / 1e6


KEYS = ("STCGAN_TC_DBGMODE", "STCGAN_TC_STAGES", "STCGAN_TC_BN256", "STCGAN_TC_BN256_STAGES", "STCGAN_TC_MT", "STCGAN_TC_BN256_AUTO",
        "STCGAN_TC_DUAL")
envs = [{}, dict(STCGAN_TC_DUAL="1"), dict(STCGAN_TC_DUAL="1", STCGAN_TC_MT="1", STCGAN_TC_BN256_AUTO="0")]
print(f"{'shape':34s} " + " ".join(f"{','.join(k[10:] + '=' + v for k, v in e.items()) or 'default':>16s}" for e in envs) + "   (us, TF/s-equivalent)")
for sh in SHAPES:
    cells = []
    for e in envs:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(e)
        us, tf = run(sh)
        cells.append(f"{us:8.1f} {tf:7.0f}")
    for k in KEYS:
        os.environ.pop(k, None)
    print(f"{sh[8]:34s} " + " ".join(cells))
