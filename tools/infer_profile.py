"""One G1 -> G2 inference step at BASELINE configs[3] (64 x 480 x 640), for an ncu launch list (NVTX range `profiled`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stcgan_b200 as S
import stcgan_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(O.REFERENCE_SEED)
G1, G2 = S.UnetGenerator(3, 1).to(dev).eval(), S.UnetGenerator(4, 3).to(dev).eval()
x = O.make_istd_batch(8, 480, 640, seed=42)[0].contiguous().repeat(B // 8, 1, 1, 1).contiguous().to(dev)
for i in range(3):
    if i == 2:
        torch.cuda.nvtx.range_push("profiled")
    S.infer(G1, G2, x)
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.nvtx.range_pop()
print("ok")
