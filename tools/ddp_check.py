"""2-rank data-parallel check on GPUs: both ranks must hold identical weights after steps; losses differ per shard."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shadow-removal-istd_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
import stcgan_b200 as S, stcgan_oracle as O
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(O.REFERENCE_SEED)
nets = dict(G1=S.UnetGenerator(3, 1), G2=S.UnetGenerator(4, 3), D1=S.NLayerDiscriminator(4), D2=S.NLayerDiscriminator(7))
for n in nets.values(): n.to(dev).train()
eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"], process_group=dist.group.WORLD)
x, m, y = (t.contiguous().to(dev) for t in O.make_istd_batch(4, 256, 256, seed=100 + rank))
use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
if use_graph:
    eng.capture(x, m, y, warmup=2)
for i in range(3):
    eng.replay() if use_graph else eng.train_step(x, m, y)
torch.cuda.synchronize()
flat = torch.cat([p.detach().reshape(-1) for n in nets.values() for p in n.parameters()])
ref = flat.clone(); dist.broadcast(ref, 0)
diff = (flat - ref).abs().max().item()
print(f"rank {rank}: graph={use_graph} losses {eng.loss_dict()['G_loss']:.4f} max |param - rank0 param| = {diff:.3e}", flush=True)
assert diff == 0.0, "replicas diverged"
dist.destroy_process_group()
