"""Run a few EAGER train steps (B=16, 256x256, bf16) -- the command profiled with ncu for the launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stcgan_b200 as S
import stcgan_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(O.REFERENCE_SEED)
nets = dict(G1=S.UnetGenerator(3, 1), G2=S.UnetGenerator(4, 3), D1=S.NLayerDiscriminator(4), D2=S.NLayerDiscriminator(7))
for n in nets.values():
    n.to(dev).train()
eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"])
x, m, y = (t.contiguous().to(dev) for t in O.make_istd_batch(batch, 256, 256))
for i in range(steps):
    S._lib.launch_count_reset()
    last = i == steps - 1
    if last:                       # ncu --nvtx --nvtx-include "profiled_step/" captures exactly this step
        torch.cuda.nvtx.range_push("profiled_step")
    eng.train_step(x, m, y)
    torch.cuda.synchronize()
    if last:
        torch.cuda.nvtx.range_pop()
    print("step", i, "launches", S._lib.launch_count(), eng.loss_dict())
