#!/usr/bin/env python
"""ST-CGAN hot-path benchmark (BASELINE.json metric: train images/sec at 256x256, 16 images per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train|train512|infer]

One "step" = one full ST-CGAN train step (src/cgan.py:274-351, VisualLoss off): G1,G2 forward, D1,D2 forward x4
each, both backward phases, two Adam updates, on one synthetic ISTD-shaped batch of 16 images per GPU.
Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM (CUDA-graph replay);
`e2e` = the same through the public API with HOST (pinned) inputs copied in and the losses read back every step.
`--impl reference` times the CPU restatement of the reference (oracle port: the reference is pure Python/torch and
cannot travel to the GPU box) on the SAME workload (16 images per step, --steps / --warmup honoured; a wall-clock cap
shortens the run only if the host is very slow, and the line says what ran), on all host threads.
`--workload train512` = BASELINE configs[4] (512x512, 32 images per GPU), `--workload infer` = configs[3].
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "shadow-removal-istd_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BATCH_PER_GPU, H, W = 16, 256, 256
METRIC = "stcgan_train_images_per_sec_256x256_b16_per_gpu"
# workload -> (images per GPU, H, W, metric, BASELINE.json config it restates)
TRAIN_WORKLOADS = {
    "train": (16, 256, 256, METRIC, "configs[1]/[2]"),
    "train512": (32, 512, 512, "stcgan_train_images_per_sec_512x512_b32_per_gpu", "configs[4]"),
}
CPU_TIME_CAP_S = 240.0      # wall-clock bound of the CPU arm (the default --steps 20 --warmup 5 run stays below it)


def train_config(workload, world):
    """The `config` object of a train line -- shared verbatim by the b200 arm and the reference arm."""
    import stcgan_oracle as O
    b, h, w, _, which = TRAIN_WORKLOADS[workload]
    return {"workload": f"ST-CGAN full train step {h}x{w}, {b} images/GPU (BASELINE {which}; cgan.py:274-351, VisualLoss off, "
                        "MSE adversarial loss as executed, Adam beta=(0.5,0.999))",
            "global_batch": b * world, "parallelism": f"dp{world}",
            "algorithmic_gflop_per_image": O.train_step_flops(h, w) / 1e9}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_oracle_rate(batch, steps, warmup=1, h=H, w=W, cap_s=CPU_TIME_CAP_S):
    """images/s of the CPU oracle (port of the reference's train step, src/cgan.py:274-351) on all host threads.
    Runs `warmup` untimed + `steps` timed steps on `batch` images; stops early (after >= 1 timed step) only if the
    wall-clock cap would be exceeded.  Returns (images/s, threads, timed steps run, warm-up steps run)."""
    import stcgan_oracle as O
    torch.set_num_threads(os.cpu_count())
    states = O.build_all_states()
    tr = O.OracleTrainer(states)
    x, m, y = O.make_istd_batch(batch, h, w, seed=42)
    t_start = time.perf_counter()
    w_run = 0
    for _ in range(max(warmup, 1)):
        tr.train_step(x, m, y)
        w_run += 1
        per = (time.perf_counter() - t_start) / w_run
        if (w_run + 1 + steps) * per > cap_s and w_run >= 1:       # a slow host: keep the time for timed steps
            break
    done, t0 = 0, time.perf_counter()
    for _ in range(steps):
        tr.train_step(x, m, y)
        done += 1
        if time.perf_counter() - t_start + (time.perf_counter() - t0) / done > cap_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, torch.get_num_threads(), done, w_run


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path (oracle port, kind "port") on the stated
    workload -- the full per-GPU batch per step, --steps / --warmup honoured -- on all host threads.  Rank 0 only."""
    if rank != 0:
        return
    wl = args.workload if args.workload in TRAIN_WORKLOADS else "train"
    b, h, w, metric, _ = TRAIN_WORKLOADS[wl]
    rate, cores, steps, warm = cpu_oracle_rate(b, max(1, args.steps), max(1, args.warmup), h, w)
    sample = (f"{steps} timed + {warm} warm-up train steps on the full {b}-image {h}x{w} batch (oracle port of src/cgan.py:274-351 on "
              "torch CPU ops, fp32; the reference is pure Python and does not travel to the GPU box)")
    if steps != args.steps:
        sample += f"; stopped early by the {CPU_TIME_CAP_S:.0f} s wall-clock cap (asked for {args.steps} steps)"
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * b / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(wl, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cudnn_baseline(dev, batch, h, w, steps=10, warmup=3):
    """"The kernel to beat" (BASELINE.md section 5): the reference's train step executed by STOCK torch on the
    same GPU -- ATen -> cuDNN convolutions / batch norm, torch.optim.Adam -- in the two forms a user of the reference could
    run today: fp32 as written (cuDNN may use TF32, torch's default), and bf16 autocast + channels_last.  The arithmetic is
    the oracle's functional restatement of the reference modules (the modules themselves do not exist on the GPU box); it is
    a reported baseline, never part of the product path.  cudnn.benchmark is ON (the baseline's best case; the reference
    itself runs with benchmark off / deterministic on, src/main.py:43-45, 90-91)."""
    import stcgan_oracle as O
    out = {}
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        for name in ("fp32", "bf16_autocast_channels_last"):
            cl = name != "fp32"
            states = {n: {k: (v.to(dev).contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v.to(dev))
                          for k, v in sd.items()} for n, sd in O.build_all_states().items()}
            tr = O.OracleTrainer(states)
            x, m, y = (t.to(dev) for t in O.make_istd_batch(batch, h, w, seed=42))
            if cl:
                x, m, y = (t.contiguous(memory_format=torch.channels_last) for t in (x, m, y))

            def step():
                if cl:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        tr.train_step(x, m, y)
                else:
                    tr.train_step(x, m, y)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
            out[name] = {"value": batch * steps / dt, "unit": "images/s", "ms_per_step": 1e3 * dt / steps, "steps": steps}
            del tr, states
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = prev
    out["what"] = ("reference train step through stock torch " + torch.__version__ + " / cuDNN "
                   + str(torch.backends.cudnn.version()) + " on this GPU (eager, cudnn.benchmark on); same batch and image size")
    return out


def instrumented_breakdown(eng, x, m, y):
    """One eager step with CUDA events around every launch of the convolution family; returns per-family
    (seconds, algorithmic flops, launches).  A long device-side sleep is queued first so that the host runs ahead and the
    event intervals contain kernel time only (measured live, on the launching stream)."""
    from stcgan_b200 import ops
    from stcgan_b200._lib import BACKEND_TC
    rec = []

    def conv_flops(geom, xx, wp, nout, oh, ow, **k):
        return 2.0 * xx.shape[0] * oh * ow * nout * xx.shape[3] * (4 if geom == 3 else 16)

    def wgrad_flops(geom, s, l, g, **k):
        return 2.0 * s.shape[0] * s.shape[1] * s.shape[2] * s.shape[3] * l.shape[3] * 16

    def thinconv_flops(t, stride, wthin, nout, oh, ow, **k):
        return 2.0 * t.shape[0] * oh * ow * nout * 128          # K padded to 16 taps x 8 channels

    def thinwgrad_flops(t, stride, thin_c, f, g, *a, **k):
        return 2.0 * f.shape[0] * f.shape[1] * f.shape[2] * f.shape[3] * 128

    def bn_apply_bytes(yy, *a, **k):
        out1, out2 = a[11], (a[13] if len(a) > 13 else k.get("out2"))
        return float(yy.numel() * 2 + out1.numel() * 2 + (0 if out2 is None else out2.numel() * 2))

    def bn_bwd_bytes(yy, ss, mi, gamma, training, g1, act1, g2, act2, acc, dy, *a, **k):
        e = dy.numel() * 2
        g2b = 0 if g2 is None else e
        return float((2 * e + g2b) * (2 if ss is not None else 1) + e)     # reduce pass + apply pass (+ dy)

    big = lambda yy, *a, **k: "big" if yy.numel() * 2 >= 16 * 2 ** 20 else "small"
    table = {
        "bn_fused_apply": (bn_apply_bytes, lambda *a, **k: "bn_fwd_" + big(*a, **k)),
        "bn_act_bwd": (bn_bwd_bytes, lambda *a, **k: "bn_bwd_" + big(*a, **k)),
        "tapconv": (conv_flops, lambda *a, **k: "conv_tc" if k.get("backend") == BACKEND_TC else "conv_ffma"),
        "tapwgrad": (wgrad_flops, lambda *a, **k: "wgrad_tc" if k.get("backend") == BACKEND_TC else "wgrad_ffma"),
        "tapconv_thin_n": (conv_flops, lambda *a, **k: "conv_tc_thin"),
        "thinconv": (thinconv_flops, lambda *a, **k: "conv_tc_thin"),
        "thinwgrad": (thinwgrad_flops, lambda *a, **k: "wgrad_tc_thin"),
    }
    saved = {}

    def timed(fn, flops_of, name_of):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            rec.append((name_of(*a, **k), flops_of(*a, **k), e0, e1))
            return out
        return wrapper

    for name, (ff, nf) in table.items():
        saved[name] = getattr(ops, name)
        setattr(ops, name, timed(saved[name], ff, nf))
    # optimiser launches: 28 B per parameter (p, m, v read + written, g read) + 4 B of refreshed bf16 copies
    adam_saved = []
    for opt in (eng.optim_D, eng.optim_G):
        for meth in ("step", "step_partial"):
            fn = getattr(opt, meth)
            adam_saved.append((opt, meth, fn))
            if meth == "step":
                nb = lambda *a, _o=opt, **k: 32.0 * sum(p.numel() for g in _o.param_groups for p in g["params"])
            else:
                nb = lambda params, *a, **k: 32.0 * sum(p.numel() for p in params)
            setattr(opt, meth, timed(fn, nb, lambda *a, **k: "adam"))
    # the instrumented step runs every kernel on ONE stream (the production step forks the weight-gradient kernels onto a
    # side stream, where per-launch event intervals on the main stream would not bracket them)
    side = {k: rt.side_stream for k, rt in eng.rt.items()}
    lanes, hi = eng.lanes.streams, eng.hi_stream
    eng.lanes.streams, eng.hi_stream = [], None
    for rt in eng.rt.values():
        rt.side_stream = None
    try:
        torch.cuda.synchronize()
        torch.cuda._sleep(int(3e9))          # ~1.5 s of device time: the whole eager step queues up behind it
        e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_all0.record()
        eng.train_step(x, m, y)
        e_all1.record()
        torch.cuda.synchronize()
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
        for opt, meth, fn in adam_saved:
            setattr(opt, meth, fn)
        for k, rt in eng.rt.items():
            rt.side_stream = side[k]
        eng.lanes.streams, eng.hi_stream = lanes, hi
    fam = {}
    for name, fl, e0, e1 in rec:
        t, f, c = fam.get(name, (0.0, 0.0, 0))
        fam[name] = (t + e0.elapsed_time(e1) * 1e-3, f + fl, c + 1)
    return fam, e_all0.elapsed_time(e_all1) * 1e-3


def ncu_traffic():
    """DRAM traffic of the dominant kernel family from the committed ncu capture of this build (written by
    tools/ncu_traffic.py from `ncu --set full` of one train step; bench.py itself never runs under a profiler)."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def run_b200(args, rank, world, local_rank):
    import stcgan_b200 as S
    import stcgan_oracle as O
    BATCH_PER_GPU, H, W, METRIC, _ = TRAIN_WORKLOADS[args.workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    torch.manual_seed(O.REFERENCE_SEED)
    nets = dict(G1=S.UnetGenerator(3, 1), G2=S.UnetGenerator(4, 3), D1=S.NLayerDiscriminator(4), D2=S.NLayerDiscriminator(7))
    for n in nets.values():
        n.to(dev).train()
    eng = S.STCGANEngine(nets["G1"], nets["G2"], nets["D1"], nets["D2"], S.TrainConfig(), process_group=pg)
    xs, ms, ys = O.make_istd_batch(BATCH_PER_GPU, H, W, seed=42 + rank)       # distinct shard per rank
    host = [t.contiguous().pin_memory() for t in (xs, ms, ys)]
    x, m, y = (t.to(dev) for t in host)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    eng.capture(x, m, y, warmup=max(args.warmup, 3))
    launches_per_step = eng.graph_launches
    for _ in range(max(args.warmup, 3)):
        eng.replay()
    # ---- device-resident throughput -------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.replay()
    e1.record()
    barrier()
    dt = e0.elapsed_time(e1) * 1e-3
    # ---- end to end: pinned host inputs in, losses out, every step --------------------------------------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    loss_host = None
    for _ in range(args.steps):
        losses = eng.replay(host[0], host[1], host[2])          # H2D copies of x, m, y inside
        loss_host = losses.cpu()                                  # D2H read of the step's losses (synchronises)
    e3.record()
    barrier()
    dt_e2e_sync = e2.elapsed_time(e3) * 1e-3
    # ---- the same, pipelined (STCGANEngine.replay_async): batch i+1's H2D copy runs under batch i's compute, the losses of
    # every step are read on the host one step later (the last one by flush()) -------------------------------------------
    eng.replay_async(*host); eng.flush()
    barrier()
    e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e6.record()
    got = 0
    for _ in range(args.steps):
        prev = eng.replay_async(*host)
        if prev is not None:
            loss_host = prev
            got += 1
    loss_host = eng.flush()
    got += 1
    e7.record()
    barrier()
    dt_e2e = e6.elapsed_time(e7) * 1e-3
    # ---- end to end with uint8 host images (the dataset's uint8 -> float transform moved onto the GPU) ----
    def to_u8(t):
        return ((t * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    host8 = [to_u8(t) for t in (xs, ms, ys)]
    eng.replay_u8(*host8)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(args.steps):
        loss_host8 = eng.replay_u8(*host8).cpu()
    e5.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dt_u8 = e4.elapsed_time(e5) * 1e-3
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt, dt_e2e, dt_u8, dt_e2e_sync], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, dt_u8, dt_e2e_sync = t.tolist()
    # the instrumented eager step contains the gradient all-reduces: every rank must run it
    fam, eager_s = instrumented_breakdown(eng, x, m, y)
    if world > 1:
        # graphs that hold NCCL work must be gone before destroy_process_group() (main() calls it after this returns)
        eng.release_graphs()
        torch.cuda.synchronize()
    if rank != 0:
        return
    peaks = measured_peaks()
    images = BATCH_PER_GPU * world * args.steps
    value, e2e_value = images / dt, images / dt_e2e
    flops_step = O.train_step_flops(H, W) * BATCH_PER_GPU
    traffic = ncu_traffic() if args.workload == "train" else None
    # ---- roofline of the dominant kernel family, timed live with CUDA events ---------------------------------------
    tc_t = sum(v[0] for k, v in fam.items() if k in ("conv_tc", "wgrad_tc"))
    tc_f = sum(v[1] for k, v in fam.items() if k in ("conv_tc", "wgrad_tc"))
    tc_n = sum(v[2] for k, v in fam.items() if k in ("conv_tc", "wgrad_tc"))
    achieved = tc_f / tc_t / 1e12 if tc_t > 0 else 0.0
    roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tf_sustained"],
            # DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the same kernel family, per launch (mean over the
            # launches of one train step), from the committed `ncu --set full` capture of this build -- or null
            "traffic": (traffic or {}).get("dominant_family_bytes_per_launch"),
            "traffic_note": (traffic or {}).get("note", "no ncu capture of this build committed (profiles/r02_ncu_traffic.json)"),
            "kernel": "tapgemm_tc_kernel + tapwgrad_tc_kernel (all full-width tcgen05 conv launches of one train step)",
            "eager_step_seconds": eager_s,
            "launches": tc_n, "flops_per_step": tc_f, "seconds_per_step": tc_t, "peak_source": peaks["source"] + ", sustained bf16",
            "families": {k: {"s": v[0], "flops": v[1], "launches": v[2]} for k, v in fam.items()
                         if k.startswith(("conv_", "wgrad_"))},
            "whole_step_tflops": flops_step / (dt / args.steps) / 1e12}
    # HBM-bound families of the same instrumented step (algorithmic bytes / CUDA-event time; "big" = tensors >= 16 MB,
    # i.e. launches long enough for the rate to mean something -- the small ones are latency-bound and L2-resident)
    hb = {k: v for k, v in fam.items() if k.startswith(("bn_", "adam"))}
    sel = list(hb.values())            # ALL BatchNorm launches (small, latency-bound ones included) + Adam
    hbm_t, hbm_b = sum(v[0] for v in sel), sum(v[1] for v in sel)
    bn_sel = [v for k, v in hb.items() if k.startswith("bn_")]
    bn_t, bn_b = sum(v[0] for v in bn_sel), sum(v[1] for v in bn_sel)
    roof_hbm = {"bound": "hbm", "achieved": hbm_b / hbm_t / 1e9 if hbm_t > 0 else 0.0, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": (hbm_b / hbm_t / 1e9 / peaks["hbm"]) if hbm_t > 0 else 0.0, "traffic": None,
                "kernel": "every bn_fused_apply / bn_bwd_reduce / bn_bwd_apply launch + adam_kernel of one train step "
                          "(algorithmic bytes / CUDA-event time, single-stream instrumented step)",
                "batchnorm_only": {"GBps": bn_b / bn_t / 1e9 if bn_t > 0 else 0.0, "frac": (bn_b / bn_t / 1e9 / peaks["hbm"]) if bn_t > 0 else 0.0,
                                   "seconds_per_step": bn_t, "launches": sum(v[2] for v in bn_sel)},
                "peak_source": peaks["source"],
                "families": {k: {"s": v[0], "bytes": v[1], "launches": v[2], "GBps": v[1] / v[0] / 1e9 if v[0] > 0 else 0.0}
                             for k, v in hb.items()}}
    for v in loss_host[:6].tolist():
        if not (v == v and abs(v) < 1e6):
            raise SystemExit(f"bench.py: non-finite / diverged loss in the timed run: {loss_host.tolist()}")
    cpu = cudnn = None
    if world == 1:
        # the CPU arm on the SAME workload (full per-GPU batch per step), bounded to ~10-30 s: 1 warm-up + 2 timed steps
        cb, ch, cw = (BATCH_PER_GPU, H, W) if args.workload == "train" else (4, H, W)
        rate, cores, done, warm = cpu_oracle_rate(cb, 2 if args.workload == "train" else 1, 1, ch, cw, cap_s=60.0)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{done} timed + {warm} warm-up train steps on {cb} of the {BATCH_PER_GPU} images per step at {ch}x{cw} "
                         "(oracle port of src/cgan.py:274-351, torch CPU fp32, all host threads)"}
        if not args.no_cudnn_baseline:
            eng.release_graphs()
            del eng, nets
            torch.cuda.empty_cache()
            try:
                cudnn = cudnn_baseline(dev, BATCH_PER_GPU, H, W)
            except Exception as e:      # a reported baseline must never cost the measured line
                cudnn = {"error": f"{type(e).__name__}: {e}"[:300]}
    cfg = train_config(args.workload, world)
    cfg.update({"precision": "bf16 activations/weights, fp32 master weights + Adam + accumulation",
                "l2": "per-step working set (>2 GB of activations and weights) exceeds the 126 MB L2; no explicit flush",
                "cuda_graph": True,
                "streams": "generator chain + four discriminator chains on five streams, weight-gradient kernels on side streams "
                           "(fork/join inside the one captured graph; gradient all-reduces inside the graph at N > 1)"})
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": sum(t.numel() * 4 for t in host),
                "d2h_bytes_per_step": int(loss_host.numel() * 4), "ms_per_step": 1e3 * dt_e2e / args.steps,
                "api": "STCGANEngine.replay_async(x, m, y) on pinned float32 host batches: every step's inputs are copied H2D "
                       "(under the previous step's compute) and every step's losses are read on the host (one step later; "
                       "the last by flush(), inside the timed region)", "loss_reads": got},
        "e2e_sync": {"value": images / dt_e2e_sync, "unit": "images/s", "ms_per_step": 1e3 * dt_e2e_sync / args.steps,
                     "note": "STCGANEngine.replay + losses.cpu(): copy, step and read strictly serialised"},
        "e2e_u8": {"value": images / dt_u8, "unit": "images/s", "h2d_bytes_per_step": sum(t.numel() for t in host8),
                   "d2h_bytes_per_step": int(loss_host.numel() * 4), "ms_per_step": 1e3 * dt_u8 / args.steps,
                   "note": "host ships decoded uint8 HWC images; uint8 -> [-1,1] float CHW on the GPU (bit-exact with the dataset code)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
        "losses_last_step": dict(zip(S.engine.SLOTS, [float(v) for v in loss_host[:6]])),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if cudnn is not None:
        line["cudnn_baseline"] = cudnn
    print(json.dumps(line), flush=True)


def run_infer(args, rank, world, local_rank):
    """BASELINE configs[3]: G1 -> G2 inference at native ISTD resolution 640x480, batch 64, bf16, eval-mode BatchNorm
    (src/cgan.py:437-446), uint8 quantisation on the GPU.  One replica per GPU (the path has no exchange step)."""
    import stcgan_b200 as S
    import stcgan_oracle as O
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, HH, WW = 64, 480, 640
    torch.manual_seed(O.REFERENCE_SEED)
    G1, G2 = S.UnetGenerator(3, 1).to(dev).eval(), S.UnetGenerator(4, 3).to(dev).eval()
    xs = O.make_istd_batch(8, HH, WW, seed=42 + rank)[0].contiguous()
    host = xs.repeat(B // 8, 1, 1, 1).contiguous().pin_memory()
    x = host.to(dev)
    for _ in range(max(args.warmup, 3)):
        S.infer(G1, G2, x)
    torch.cuda.synchronize()
    S._lib.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        S.infer(G1, G2, x)
    e1.record()
    torch.cuda.synchronize()
    launches = S._lib.launch_count()
    dt = e0.elapsed_time(e1) * 1e-3
    # end to end through the public API (stcgan_b200.infer_u8): decoded uint8 HWC images in pinned host memory in, uint8 HWC
    # mask / shadow-free images in pinned host memory out, every step; the uint8 <-> float transforms run on the GPU
    host8 = ((xs * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).repeat(B // 8, 1, 1, 1).contiguous().pin_memory()
    out_m = torch.empty((B, HH, WW, 1), dtype=torch.uint8).pin_memory()
    out_y = torch.empty((B, HH, WW, 3), dtype=torch.uint8).pin_memory()
    S.infer_u8(G1, G2, host8, out_m, out_y)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        S.infer_u8(G1, G2, host8, out_m, out_y)
        torch.cuda.current_stream().synchronize()            # the step's results are on the host
    e3.record()
    torch.cuda.synchronize()
    dt2 = e2.elapsed_time(e3) * 1e-3
    m_host, y_host = out_m, out_y
    # the same through stcgan_b200.InferencePipeline: every step's images go host -> device and every step's results device ->
    # host, on copy streams underneath the neighbouring steps' compute (results are handed out one step later; flush() inside)
    pipe = S.InferencePipeline(G1, G2)
    pipe.submit(host8); pipe.submit(host8); pipe.flush()
    torch.cuda.synchronize()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    got = 0
    for _ in range(args.steps):
        if pipe.submit(host8) is not None:
            got += 1
    res = pipe.flush()
    got += 1
    e5.record()
    torch.cuda.synchronize()
    dt3 = e4.elapsed_time(e5) * 1e-3
    assert res is not None and res[1].shape == y_host.shape
    if rank != 0:
        return
    peaks = measured_peaks()
    flops = O.inference_flops(HH, WW) * B
    line = {"metric": "stcgan_infer_images_per_sec_480x640_b64", "value": B * world * args.steps / dt, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "G1->G2 inference 480x640, batch 64, eval-mode BN, uint8 outputs (BASELINE configs[3])",
                       "l2": "activations of one step (GBs) exceed the 126 MB L2", "parallelism": f"replicas x{world}",
                       "algorithmic_gflop_per_image": O.inference_flops(HH, WW) / 1e9},
            "e2e": {"value": B * world * args.steps / dt3, "unit": "images/s", "h2d_bytes_per_step": host8.numel(),
                    "d2h_bytes_per_step": int(m_host.numel() + y_host.numel()), "ms_per_step": 1e3 * dt3 / args.steps,
                    "api": "InferencePipeline.submit(uint8 HWC pinned batch) -> uint8 HWC results in pinned memory, one step later "
                           "(H2D / D2H on copy streams under the neighbouring steps' compute; the last by flush(), timed)",
                    "result_reads": got},
            "e2e_sync": {"value": B * world * args.steps / dt2, "unit": "images/s", "ms_per_step": 1e3 * dt2 / args.steps,
                         "note": "infer_u8 + stream synchronise per step: copy in, compute, copy out strictly serialised"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": flops / (dt / args.steps) / 1e12, "peak": peaks["tf_sustained"],
                         "unit": "TFLOP/s", "frac": flops / (dt / args.steps) / 1e12 / peaks["tf_sustained"], "traffic": None,
                         "kernel": "whole inference step (all kernels)", "peak_source": peaks["source"]}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "train512", "infer"])
    ap.add_argument("--no-cudnn-baseline", action="store_true", help="skip the stock torch/cuDNN arm (N = 1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus and args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    if args.workload == "infer":
        run_infer(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
