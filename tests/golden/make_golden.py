"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

It imports `src.networks`, `src.loss`, `src.utils` from the reference, builds the four
networks with the reference constructors under the reference seed (src/main.py:239),
runs the restated step of src/cgan.py:274-351 (VisualLoss off) in float64 and float32 and
the inference path of src/cgan.py:437-442, and stores *samples* (strided element picks),
norms and scalars -- not whole tensors -- so the fixtures stay small:

    stcgan_step_b2_256.npz     forward outputs, losses, gradients, post-Adam weights, BN buffers
    stcgan_infer_480x640.npz   eval-mode G1->G2 outputs and their float2uint quantisation
    float2uint_vectors.npz     adversarial float inputs and the reference's uint8 outputs
    weights_fingerprint.npz    per-tensor (sum, sum|.|) of the seeded default-init weights

The reference has no golden vectors of its own; these are outputs of the
reference itself, which is the strongest pin available.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stcgan_oracle as O                      # noqa: E402  (input generator + hyper-parameters only)
import pin_against_reference as P              # noqa: E402  (reference loaders + restated step on ref modules)

STRIDE = 997   # prime stride for element sampling


def sample(t):
    f = t.detach().reshape(-1).double().numpy()
    return f[::STRIDE].copy() if f.size > 4096 else f.copy()


def main():
    networks, loss, utils = P.load_reference()
    torch.set_num_threads(os.cpu_count())
    hp = O.HyperParams()
    out_dir = HERE

    # ---- weights fingerprint ------------------------------------------------
    nets = P.build_reference_nets(networks)
    fp = {}
    for n, net in nets.items():
        for k, v in net.state_dict().items():
            if v.is_floating_point():
                fp[f"{n}/{k}"] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
    np.savez_compressed(os.path.join(out_dir, "weights_fingerprint.npz"), **fp)

    # ---- train step, float64 (rigorous) and float32 (as executed) --------------
    gold = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        nets = P.build_reference_nets(networks)
        for n in nets:
            nets[n].to(dtype).train()
        optim_G = torch.optim.Adam(list(nets["G1"].parameters()) + list(nets["G2"].parameters()),
                                   lr=hp.lr_G, betas=(hp.beta1, hp.beta2))
        optim_D = torch.optim.Adam(list(nets["D1"].parameters()) + list(nets["D2"].parameters()),
                                   lr=hp.lr_D, betas=(hp.beta1, hp.beta2))
        adv = loss.AdversarialLoss(ls=hp.ls, rel=hp.rel, avg=hp.avg).to(dtype)
        dl = loss.DataLoss()
        x, m, y = (t.to(dtype) for t in O.make_istd_batch(2, 256, 256, seed=42))
        r = P.reference_train_step(nets, adv, dl, optim_G, optim_D, x, m, y, hp)
        for k in ("m_pred", "y_pred", "C1_fake_Gphase", "C2_fake_Gphase"):
            gold[f"{tag}/{k}/sample"] = sample(r[k])
            gold[f"{tag}/{k}/norm"] = np.array(r[k].double().norm().item())
        for k in ("D1_loss", "D2_loss", "D_loss", "G1_loss", "G2_loss", "data1_loss", "data2_loss", "G_loss"):
            gold[f"{tag}/{k}"] = np.array(r[k].double().item())
        for grp in ("grads_D", "grads_G"):
            for n in r[grp]:
                names = [k for k, _ in nets[n].named_parameters()]
                for k, g in zip(names, r[grp][n]):
                    gold[f"{tag}/grad/{n}/{k}/sample"] = sample(g)
                    gold[f"{tag}/grad/{n}/{k}/norm"] = np.array(g.double().norm().item())
        for n in nets:
            for k, v in nets[n].state_dict().items():
                if "running" in k or "num_batches" in k:
                    gold[f"{tag}/post/{n}/{k}"] = v.double().numpy() if v.is_floating_point() else v.numpy()
                elif tag == "f64":
                    gold[f"{tag}/post/{n}/{k}/sample"] = sample(v)
    gold["stride"] = np.array(STRIDE)
    np.savez_compressed(os.path.join(out_dir, "stcgan_step_b2_256.npz"), **gold)

    # ---- adversarial-loss branches on a fixed logit pair ------------------------
    g = torch.Generator().manual_seed(11)
    cr = torch.randn(3, 1, 30, 30, generator=g); cf = torch.randn(3, 1, 30, 30, generator=g)
    lossgold = {"C_real": cr.numpy(), "C_fake": cf.numpy()}
    for ls in (False, True):
        for rel in (False, True):
            for avg in (False, True):
                for d in (False, True):
                    a = loss.AdversarialLoss(ls=ls, rel=rel, avg=avg)
                    crg, cfg = cr.clone().requires_grad_(True), cf.clone().requires_grad_(True)
                    v = a(crg, cfg, D_loss=d)
                    v.backward()
                    key = f"ls{int(ls)}_rel{int(rel)}_avg{int(avg)}_D{int(d)}"
                    lossgold[key] = np.array(v.item(), dtype=np.float32)
                    lossgold[key + "/dreal"] = (crg.grad if crg.grad is not None else torch.zeros_like(cr)).numpy()
                    lossgold[key + "/dfake"] = (cfg.grad if cfg.grad is not None else torch.zeros_like(cf)).numpy()
    np.savez_compressed(os.path.join(out_dir, "adversarial_loss.npz"), **lossgold)

    # ---- inference at native ISTD size -----------------------------------------
    inf = {}
    nets = P.build_reference_nets(networks)
    nets["G1"].eval(); nets["G2"].eval()
    x = O.make_istd_batch(1, 480, 640, seed=5)[0]
    with torch.no_grad():
        mp = nets["G1"](x); yp = nets["G2"](torch.cat((x, mp), 1))
    m_np = mp.numpy() * 0.5 + 0.5; y_np = yp.numpy() * 0.5 + 0.5          # cgan.py:441-442
    m_u8 = utils.float2uint(m_np[0].transpose(1, 2, 0)); y_u8 = utils.float2uint(y_np[0].transpose(1, 2, 0))
    inf["m_pred/sample"] = sample(mp); inf["y_pred/sample"] = sample(yp)
    inf["m_pred/norm"] = np.array(mp.double().norm().item()); inf["y_pred/norm"] = np.array(yp.double().norm().item())
    inf["m_u8/sample"] = m_u8.reshape(-1)[::STRIDE].copy(); inf["y_u8/sample"] = y_u8.reshape(-1)[::STRIDE].copy()
    inf["m_u8/hist"] = np.bincount(m_u8.reshape(-1), minlength=256)
    inf["y_u8/hist"] = np.bincount(y_u8.reshape(-1), minlength=256)
    np.savez_compressed(os.path.join(out_dir, "stcgan_infer_480x640.npz"), **inf)

    # ---- float2uint known-answer vectors -----------------------------------------
    rs = np.random.RandomState(0)
    grid = (np.arange(0, 257, dtype=np.float32) / 255.0)
    vec = np.concatenate([
        grid, np.nextafter(grid, np.float32(-1)), np.nextafter(grid, np.float32(2)),
        np.array([-1.0, -0.0, 0.0, 1e-9, 0.99999994, 1.0, 1.0000001, 2.0, 1e9], np.float32),
        rs.rand(4096).astype(np.float32) * 1.4 - 0.2,
        np.tanh(rs.randn(4096)).astype(np.float32) * 0.5 + 0.5,
    ]).astype(np.float32)
    np.savez_compressed(os.path.join(out_dir, "float2uint_vectors.npz"),
                        inputs=vec, outputs=utils.float2uint(vec),
                        pre=np.tanh(rs.randn(4096)).astype(np.float32))
    print("wrote fixtures to", out_dir)
    for f in sorted(os.listdir(out_dir)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(out_dir, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
