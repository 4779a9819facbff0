"""N>1 host logic on CPU: world_size-2 gloo process groups.

1. `GradientSync` (the exchange the engine runs between its CUDA-graph segments): async SUM all-reduce of named flat
   buffers, blocking semantics, call order.
2. The data-parallel contract itself: every rank runs its own shard with rank-local BatchNorm statistics,
   gradients are summed and scaled by 1/world -- this equals a single process that runs each shard through a SEPARATE
   forward and averages the gradients (nn.DataParallel's per-replica semantics, src/cgan.py:78-84).  Checked with the CPU
   oracle on a narrow (ngf=8) network so that it runs in seconds.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "shadow-removal-istd_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import stcgan_oracle as O
    from stcgan_b200.engine import GradientSync

    # ---- 1. GradientSync protocol
    bufs = {"D1": torch.full((5,), float(rank + 1)), "D2": torch.full((3,), 10.0 * (rank + 1)),
            "G2": torch.arange(4.0) * (rank + 1), "G1": torch.ones(2) * (rank + 1)}
    sync = GradientSync(lambda n: bufs[n], dist.group.WORLD)
    pending = []
    sync.reduce(("D1", "D2"), True, pending)
    assert not pending and torch.equal(bufs["D1"], torch.full((5,), 3.0)) and torch.equal(bufs["D2"], torch.full((3,), 30.0))
    sync.reduce(("G2",), False, pending)
    assert len(pending) == 1                      # still in flight: overlaps the next segment
    sync.reduce(("G1",), True, pending)
    assert not pending and torch.equal(bufs["G2"], torch.arange(4.0) * 3) and torch.equal(bufs["G1"], torch.ones(2) * 3)
    assert sync.log == [(("D1", "D2"), True), (("G2",), False), (("G1",), True)] and sync.world == 2
    # bucketed schedule of the engine's G phase: G2 async, then G1's up-conv bucket async while G2 must have landed,
    # then the rest async while the up-conv bucket must have landed, then a blocking wait for everything
    bufs.update({"G2": torch.ones(4) * (rank + 1), "G1.ups": torch.ones(3) * (rank + 1), "G1.rest": torch.ones(2) * (rank + 1)})
    pending = []
    sync.reduce(("G2",), False, pending)
    sync.reduce(("G1.ups",), False, pending, wait=("G2",))
    assert [n for n, _ in pending] == ["G1.ups"] and torch.equal(bufs["G2"], torch.ones(4) * 3)
    sync.reduce(("G1.rest",), False, pending, wait=("G1.ups",))
    assert [n for n, _ in pending] == ["G1.rest"] and torch.equal(bufs["G1.ups"], torch.ones(3) * 3)
    sync.reduce((), True, pending)
    assert not pending and torch.equal(bufs["G1.rest"], torch.ones(2) * 3)

    # ---- 2. rank-local BN + summed gradients / world == per-shard forwards averaged
    torch.manual_seed(123)
    sd = O.build_discriminator_state(4, ndf=8)
    keys = O.trainable_keys(sd)
    x, m, _ = O.make_istd_batch(2 * world, 64, 64, seed=9)
    inp = torch.cat((x, m), 1)

    def shard_grads(r):
        s = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd.items()}
        c = O.discriminator_forward(s, inp[2 * r:2 * r + 2], training=True)
        O.cal_loss(c, 1.0).backward()
        return [s[k].grad for k in keys]

    mine = shard_grads(rank)
    flat = torch.cat([g.reshape(-1) for g in mine])
    dist.all_reduce(flat)
    flat /= world
    ref = [sum(gs) / world for gs in zip(*[shard_grads(r) for r in range(world)])]
    ref_flat = torch.cat([g.reshape(-1) for g in ref])
    assert torch.allclose(flat, ref_flat, rtol=1e-5, atol=1e-7)
    # ...and it is NOT what one forward over the concatenated batch gives (global BN statistics differ): SyncBN would be wrong
    s = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd.items()}
    O.cal_loss(O.discriminator_forward(s, inp, training=True), 1.0).backward()
    glob = torch.cat([s[k].grad.reshape(-1) for k in keys])
    assert (glob - ref_flat).norm() / ref_flat.norm() > 1e-3
    # ---- 3. the engine's D-phase protocol since round 2: D1's bucket first (blocking on its own lane), then D2's
    bufs.update({"D1": torch.ones(4) * (rank + 1), "D2": torch.ones(4) * 2 * (rank + 1)})
    sync.log.clear(); pending = []
    sync.reduce(("D1",), True, pending)
    sync.reduce(("D2",), True, pending)
    assert not pending and torch.equal(bufs["D1"], torch.ones(4) * 3) and torch.equal(bufs["D2"], torch.ones(4) * 6)
    assert sync.log == [(("D1",), True), (("D2",), True)]

    # ---- 4. the single-process data-parallel oracle used by the 2-GPU parity test (tests/test_ddp_gpu.py) states the same
    # thing as real ranks: D-phase gradients of O.OracleDataParallel == all-reduced per-rank gradients / world
    torch.manual_seed(7)
    st = O.build_all_states(ngf=8, ndf=8)
    shards = [tuple(t.double() for t in O.make_istd_batch(1, 256, 256, seed=40 + r)) for r in range(world)]
    mine = O.OracleTrainer(st, dtype=torch.float64).train_step(*shards[rank], do_optim=False, keep_grads=True)
    flat = torch.cat([g.reshape(-1) for n in ("D1", "D2") for g in mine["grads_D"][n]])
    dist.all_reduce(flat)
    flat /= world
    _, g = O.OracleDataParallel(st, world, dtype=torch.float64).train_step(shards)
    want = torch.cat([t.reshape(-1) for n in ("D1", "D2") for t in g[n]])
    assert torch.allclose(flat, want, rtol=1e-9, atol=1e-12), float((flat - want).abs().max())
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_sync_and_dp_semantics(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
