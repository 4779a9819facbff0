"""Run-to-run spread of one fp32-mode engine step (same weights, same batch) with the step's concurrency on and off:
if the multi-stream schedule had a race, the spread with concurrency on would exceed the atomics-order noise seen with it off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stcgan_b200 as S
import stcgan_oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
states = O.build_all_states()
x, m, y = (t.contiguous().to(dev) for t in O.make_istd_batch(2, 256, 256, seed=7))
ref = O.OracleTrainer(states, dtype=torch.float64).train_step(*(t.double().cpu() for t in (x, m, y)))
for conc in ("1", "0"):
    os.environ["STCGAN_CONCURRENCY"] = conc
    vals = []
    for rep in range(6):
        nets = dict(G1=S.UnetGenerator(3, 1, precision=mode), G2=S.UnetGenerator(4, 3, precision=mode),
                    D1=S.NLayerDiscriminator(4, precision=mode), D2=S.NLayerDiscriminator(7, precision=mode))
        for n, mod in nets.items():
            mod.load_state_dict(states[n]); mod.to(dev).train()
        eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"])
        eng.train_step(x, m, y)
        torch.cuda.synchronize()
        L = eng.loss_dict()
        vals.append([L[k] for k in ("D1_loss", "D2_loss", "G1_loss", "G2_loss", "data1_loss", "data2_loss")])
    t = torch.tensor(vals, dtype=torch.float64)
    want = torch.tensor([float(ref[k]) for k in ("D1_loss", "D2_loss", "G1_loss", "G2_loss", "data1_loss", "data2_loss")], dtype=torch.float64)
    print(f"mode {mode} concurrency={conc}: rel spread (max-min)/mean per loss", ((t.max(0).values - t.min(0).values) / t.mean(0)).tolist())
    print(f"   mean rel deviation from the fp64 oracle", ((t.mean(0) - want) / want).tolist())
