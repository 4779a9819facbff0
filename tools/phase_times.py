"""Timeline of one eager train step with the production stream layout (lanes + side streams): start / end of every network
forward / backward pass and optimiser call relative to the start of the step, from CUDA events recorded on the stream each
call was issued on.  Usage: python tools/phase_times.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stcgan_b200 as S
import stcgan_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(O.REFERENCE_SEED)
nets = dict(G1=S.UnetGenerator(3, 1), G2=S.UnetGenerator(4, 3), D1=S.NLayerDiscriminator(4), D2=S.NLayerDiscriminator(7))
for n in nets.values():
    n.to(dev).train()
eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"])
x, m, y = (t.contiguous().to(dev) for t in O.make_istd_batch(16, 256, 256))
for _ in range(3):
    eng.train_step(x, m, y)
torch.cuda.synchronize()
rec = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        rec.append((label, e0, e1))
        return out
    setattr(obj, name, w)


for k, rt in eng.rt.items():
    wrap(rt, "forward", f"{k} forward")
    wrap(rt, "backward", f"{k} backward")
wrap(eng.optim_D, "step", "Adam D")
wrap(eng.optim_G, "step", "Adam G")
wrap(eng.optim_G, "step_partial", "Adam G (partial)")
torch.cuda._sleep(int(4e9))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
eng.train_step(x, m, y)
t1.record()
torch.cuda.synchronize()
print(f"eager step {t0.elapsed_time(t1):.3f} ms")
for label, e0, e1 in rec:
    a, b = t0.elapsed_time(e0), t0.elapsed_time(e1)
    print(f"{a:8.3f} -> {b:8.3f} ms  ({b - a:6.3f})  {label}")
