"""Debug the thin-K path: identity weights expose the im2col rows the tensor core actually saw."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200"))
import torch
from stcgan_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (n, h, w, s, border) in [(1, 8, 16, 2, 1), (2, 16, 16, 2, 1), (1, 7, 9, 1, 2), (2, 32, 32, 2, 1)]:
    hp, wp = h + 2 * border, w + 2 * border
    t = torch.zeros(n, hp, wp, 8, dtype=torch.bfloat16, device=dev)
    # value encodes (y, x, c): small integers exactly representable in bf16
    yy, xx, cc = torch.meshgrid(torch.arange(hp), torch.arange(wp), torch.arange(8), indexing="ij")
    val = ((yy * 7 + xx) % 61 + cc * 0.125).to(torch.bfloat16).to(dev)
    t[:] = val
    oh = (hp - 4) // s + 1; ow = (wp - 4) // s + 1
    wt = torch.eye(128, dtype=torch.bfloat16, device=dev)       # Wt[n][k] = delta
    out = ops.thinconv(t, s, wt, 128, oh, ow)
    torch.cuda.synchronize()
    # expected im2col
    exp = torch.zeros(n, oh, ow, 128)
    tc = t.float().cpu()
    for kh in range(4):
        for kw in range(4):
            exp[..., (kh * 4 + kw) * 8:(kh * 4 + kw) * 8 + 8] = tc[:, kh:kh + s * (oh - 1) + 1:s, kw:kw + s * (ow - 1) + 1:s, :]
    got = out.float().cpu()
    bad = (got != exp)
    print(f"case n={n} h={h} w={w} s={s}: out {tuple(got.shape)} mismatches {int(bad.sum())} / {bad.numel()}")
    if bad.any():
        rows = bad.reshape(-1, 128).any(dim=1).nonzero().flatten()[:6].tolist()
        for r in rows:
            cols = bad.reshape(-1, 128)[r].nonzero().flatten().tolist()
            print("  pixel", r, "bad cols", cols[:16], "...", len(cols))
            print("    got", got.reshape(-1, 128)[r, :16].tolist())
            print("    exp", exp.reshape(-1, 128)[r, :16].tolist())
