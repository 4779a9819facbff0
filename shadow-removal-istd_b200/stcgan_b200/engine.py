"""Hand-scheduled ST-CGAN train step and inference on the kernel runtimes (no autograd).

`STCGANEngine.train_step` executes exactly the step body of the reference trainer
(CGAN.run_epoch, src/cgan.py:274-351, VisualLoss terms off):

    D phase : C1_real = D1(x,m); m_pred = G1(x); C1_fake = D1(x, m_pred.detach());
              C2_real = D2(x,m,y); y_pred = G2(x, m_pred); C2_fake = D2(x, m_pred.detach(), y_pred.detach());
              D_loss = l2*D1_loss + l3*D2_loss; backward; Adam(D1+D2)
    G phase : the four D passes again with the UPDATED discriminators (the two `real` ones only matter for
              BatchNorm running statistics -- SURVEY appendix A -- and can be skipped behind a flag);
              G_loss = L1(m_pred,m) + l1*L1(y_pred,y) + l2*adv(C1_fake) + l3*adv(C2_fake); backward through
              D (dgrad only), G2 (its input gradient's mask channel flows into G1) and G1; Adam(G1+G2)

Gradients stay in packed tap-major layout and are consumed in place by the fused Adam; under
data parallelism (one process per GPU) the four flat gradient buffers are all-reduced with NCCL, the G2
bucket while G1's backward is still running.  The whole step can be captured into ONE CUDA graph.
"""
from __future__ import annotations

import atexit
import os
import weakref
from dataclasses import dataclass

import torch

from . import _lib, ops
from .optim import FusedAdam

SLOTS = ("D1_loss", "D2_loss", "G1_loss", "G2_loss", "data1_loss", "data2_loss", "vis1_loss", "vis2_loss")

_LIVE_ENGINES = weakref.WeakSet()


def _release_all_graphs():
    """atexit: captured graphs that contain NCCL collectives must be destroyed before the process group / the CUDA context
    go away (torch 2.11 + NCCL 2.28 otherwise hang in destroy_process_group() or at interpreter exit)."""
    for e in list(_LIVE_ENGINES):
        try:
            e.release_graphs()
        except Exception:
            pass


atexit.register(_release_all_graphs)


class _Lanes:
    """Extra streams for the D1 / D2 chains of a phase (lanes 0, 1; lanes 2, 3 take the second backward pass of each).  fork(): the lanes wait for everything issued so far on the
    current stream; lane(i): context manager that issues on lane i (or on the current stream when disabled);
    lane_wait(i): lane i additionally waits for what the current stream has issued since; join(): the current stream waits
    for both lanes."""

    def __init__(self, streams):
        self.streams = streams

    def fork(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)

    def lane(self, i):
        import contextlib
        return torch.cuda.stream(self.streams[i % len(self.streams)]) if self.streams else contextlib.nullcontext()

    def lane_wait(self, i):
        if self.streams:
            self.streams[i % len(self.streams)].wait_stream(torch.cuda.current_stream())

    def wait_lane(self, i, j):
        """lane i waits for everything issued so far on lane j"""
        if self.streams and i % len(self.streams) != j % len(self.streams):
            self.streams[i % len(self.streams)].wait_stream(self.streams[j % len(self.streams)])

    def join(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)

    def join_lanes(self, idx):
        """the current stream waits for the given lanes only"""
        if self.streams:
            cur = torch.cuda.current_stream()
            for i in idx:
                cur.wait_stream(self.streams[i % len(self.streams)])

    def record(self, i):
        """event after everything issued so far on lane i (None when the lanes are disabled: program order suffices)"""
        if not self.streams:
            return None
        ev = torch.cuda.Event()
        ev.record(self.streams[i % len(self.streams)])
        return ev

    def wait_event(self, i, ev):
        if self.streams and ev is not None:
            self.streams[i % len(self.streams)].wait_event(ev)


class GradientSync:
    """Data-parallel gradient exchange: SUM all-reduce of named flat gradient buffers over a process group (NCCL on the
    GPUs, gloo in the CPU tests); the 1/world scale is applied inside the optimiser kernel.  `reduce(names, blocking,
    pending, wait=())` launches the collectives asynchronously; `wait` names earlier reductions that must have landed before
    the caller continues, a blocking call waits for everything pending -- this is how the G2 bucket and G1's up-conv bucket
    overlap G1's backward while D's bucket gates `optim_D.step` (src/cgan.py:305)."""

    def __init__(self, buffers, process_group=None):
        self.buffers = buffers          # callable name -> tensor (buffers may be allocated lazily)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.log = []                   # (names, blocking) in call order, for tests

    def reduce(self, names, blocking, pending, wait=()):
        self.log.append((tuple(names), bool(blocking)))
        if self.world > 1:
            import torch.distributed as dist
            for n in names:
                pending.append((n, dist.all_reduce(self.buffers(n), op=dist.ReduceOp.SUM, group=self.pg, async_op=True)))
        keep = []
        for n, wk in pending:
            if blocking or n in wait:
                wk.wait()
            else:
                keep.append((n, wk))
        pending[:] = keep


@dataclass
class TrainConfig:
    """Hyper-parameters of the step.  lr / beta / lambda1..3 are the defaults of src/main.py:182-239; `ls` is what
    `args.D_loss_fn == "leastsqure"` evaluates to (always False, src/cgan.py:147); `rel` / `avg` select the RpGAN / RaGAN
    forms of AdversarialLoss (src/loss.py:88-96, 102-110; guild.yml:16-18 defaults to rel_avg).

    lambda4 / lambda5 weight the VGG19 perceptual terms vis1 / vis2 (src/cgan.py:334-348).  NOTE: the reference's own
    defaults are lambda4 = 5, lambda5 = 50 (src/main.py); they default to 0 HERE because `VisualLoss` needs pretrained
    VGG19-BN weights that this package does not ship and does not compute.  Setting either to a non-zero
    value requires passing `visual_loss=` (any differentiable torch callable `(pred, target) -> scalar`, e.g. the
    reference's `src.loss.VisualLoss().cuda()`) to STCGANEngine: its value and its gradient w.r.t. m_pred / y_pred are
    taken with torch autograd and added to the fused loss gradients before G2's / G1's backward."""
    lr_G: float = 5e-4
    lr_D: float = 1e-4
    beta1: float = 0.5
    beta2: float = 0.999
    lambda1: float = 5.0
    lambda2: float = 0.5
    lambda3: float = 0.5
    lambda4: float = 0.0
    lambda5: float = 0.0
    ls: bool = False
    rel: bool = False
    avg: bool = False
    skip_dead_real_passes: bool = False   # True drops the two G-phase `real` D passes (changes D's running stats only;
                                          # ignored for the relativistic losses, which need them)


class STCGANEngine:
    def __init__(self, G1, G2, D1, D2, cfg: TrainConfig = TrainConfig(), process_group=None, visual_loss=None):
        self.nets = dict(G1=G1, G2=G2, D1=D1, D2=D2)
        self.cfg = cfg
        if (cfg.lambda4 != 0 or cfg.lambda5 != 0) and visual_loss is None:
            raise NotImplementedError(
                "lambda4 / lambda5 weight the VGG19 perceptual loss (src/cgan.py:334-348, src/loss.py:29-56), which this "
                "package does not implement (pretrained weights are not available offline): pass visual_loss=<torch "
                "callable (pred, target) -> scalar>, e.g. the reference's VisualLoss().cuda(), or leave them at 0")
        if cfg.avg and not cfg.rel:
            raise ValueError("avg=True only has a meaning together with rel=True (src/loss.py:88-110)")
        self.visual_loss = visual_loss
        for n in self.nets.values():
            if not next(n.parameters()).is_cuda:
                raise RuntimeError("STCGANEngine needs the modules on a CUDA device (no CPU fallback)")
            n.train()
        self.rt = {k: n.runtime() for k, n in self.nets.items()}
        for rt in self.rt.values():
            rt.ensure_packed()
            rt.alloc_grads()
        self.device = self.rt["G1"].device()
        # Concurrency inside the step (STCGAN_CONCURRENCY=0 turns all of it off, STCGAN_SIDE_STREAM=0 only the first):
        #  * every network's weight-gradient kernels run on its own side stream, next to the dgrad / BatchNorm chain
        #    (nets._wgrad_async);
        #  * the D1 chain and the D2 chain of each phase run on lanes next to the generator chain on the main stream --
        #    they only meet at the loss kernels (cgan.py:281-302, 321-348 have no other cross-dependency).
        # All forks and joins are stream events, so the same code runs eagerly and under CUDA-graph capture.
        conc = os.environ.get("STCGAN_CONCURRENCY", "1") != "0"
        side = conc and os.environ.get("STCGAN_SIDE_STREAM", "1") != "0"
        mk = lambda: torch.cuda.Stream(device=self.device)
        self.side_streams = {k: (mk() if side else None) for k in ("G", "D1", "D2")}
        for k, r in self.rt.items():
            r.side_stream = self.side_streams["G" if k in ("G1", "G2") else k]
        self.lanes = _Lanes([mk(), mk(), mk(), mk()] if conc else [])
        # opt-in (STCGAN_HI_PRIORITY=1): the generator chain -- the step's critical path -- on a HIGH-priority stream.
        # Measured on B200: helps the eager step (6.65 -> 6.42 ms: G1's forward 1.37 -> 0.69 ms) but not the captured
        # graph (6.13 -> 6.25 ms, with the priority carried as a kernel launch attribute), so it is off by default
        hi = conc and os.environ.get("STCGAN_HI_PRIORITY", "0") == "1"
        self.hi_stream = torch.cuda.Stream(device=self.device, priority=-1) if hi else None
        self.optim_G = FusedAdam(list(G1.parameters()) + list(G2.parameters()), lr=cfg.lr_G, betas=(cfg.beta1, cfg.beta2))
        self.optim_D = FusedAdam(list(D1.parameters()) + list(D2.parameters()), lr=cfg.lr_D, betas=(cfg.beta1, cfg.beta2))
        self.optim_G.set_packed_grads({**self.rt["G1"].param_grad_views, **self.rt["G2"].param_grad_views})
        self.optim_D.set_packed_grads({**self.rt["D1"].param_grad_views, **self.rt["D2"].param_grad_views})
        self.optim_G.set_pack_targets(self.rt["G1"].convs + self.rt["G2"].convs)
        # optim_G is applied in three partial launches, one per gradient bucket, in the order the buckets become final:
        # G2, G1's up-conv weights ("G1.ups"), the rest of G1
        g2_ids = {id(p) for p in G2.parameters()}
        ups_ids = {id(c.weight) for c in self.rt["G1"].ups}
        deep_ids = {id(c.weight) for c in self.rt["G1"].deep_convs}
        self._g1_ups = [c.weight for c in self.rt["G1"].ups]
        self._g1_rest = [p for p in G1.parameters() if id(p) not in ups_ids]
        self._g1_deep = [c.weight for c in self.rt["G1"].deep_convs]
        self._g1_tail = [p for p in G1.parameters() if id(p) not in ups_ids and id(p) not in deep_ids]
        self.optim_G.set_block_order(lambda p: 0 if id(p) in g2_ids else (1 if id(p) in ups_ids else (2 if id(p) in deep_ids else 3)))
        self.optim_D.set_pack_targets(self.rt["D1"].convs + self.rt["D2"].convs)
        def bucket(name):                  # "G1" -> whole flat buffer, "G1.ups" / "G1.deep" / "G1.tail" / "G1.rest" -> its slices
            net, _, part = name.partition(".")
            return self.rt[net].grad_bucket(part or None)
        self.sync = GradientSync(bucket, process_group)
        self.pg, self.world = process_group, self.sync.world
        self.optim_G.grad_scale = self.optim_D.grad_scale = 1.0 / self.world
        self.losses = torch.zeros(8, dtype=torch.float32, device=self.device)
        # schedule switches (experiments / A-B measurements; the defaults are the measured-best combination)
        flag = lambda name, default: os.environ.get(name, default) != "0"
        self._early_d1 = conc and flag("STCGAN_EARLY_D1", "1")        # D1's exchange + Adam share + G-phase passes under G2's forward
        # G2's / G1.ups' / G1.deep's Adam shares underneath the halves of G1's backward: +0.5 % on one GPU, but at N > 1 the
        # HBM-bound Adam kernels slow the NCCL kernels that share G1's backward with them, the serial chain of all-reduces
        # slips and the last bucket lands late (measured at 8 GPUs: 7.42 ms per step with it, 6.98 ms without; a hybrid with only
        # G1's up-conv share moved under the encoder half: 6.70 against 6.61 ms at 2 GPUs) -> single GPU only
        self._adam_early = conc and flag("STCGAN_ADAM_EARLY", "1" if self.world == 1 else "0")
        self._overlap_real = conc and flag("STCGAN_OVERLAP_REAL", "1")  # D2's G-phase real pass next to its fake pass
        self._deep_bucket = flag("STCGAN_DEEP_BUCKET", "1")           # N > 1: G1's deep encoder gradients go on the wire early
        self._graph = None
        self._graphs = []
        self._static = None
        self.last = {}
        _LIVE_ENGINES.add(self)

    def _critical(self):
        """Context: issue on the high-priority stream (forked from / joined into the current stream)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            if self.hi_stream is None:
                yield
                return
            cur = torch.cuda.current_stream()
            self.hi_stream.wait_stream(cur)
            with torch.cuda.stream(self.hi_stream):
                yield
            cur.wait_stream(self.hi_stream)
        return ctx()

    # ------------------------------------------------------------------------------------------
    # adversarial objectives (src/loss.py:86-112) as launches of the fused loss kernel; they return the gradients
    # w.r.t. C_real / C_fake (None where the caller does not need one) and accumulate the loss value into `slot`
    # ------------------------------------------------------------------------------------------
    def _labels(self):
        cfg = self.cfg
        return (ops.KIND_BCE if cfg.ls else ops.KIND_MSE), 1.0, (-1.0 if cfg.ls else 0.0)

    def _rel_terms(self, c_real, c_fake, d_loss, lam, slot):
        """RpGAN / RaGAN objective of one discriminator, D side (d_loss) or G side: value into `slot`, returns
        (dL/dC_real, dL/dC_fake) scaled by `lam`."""
        cfg = self.cfg
        kind, real, fake = self._labels()
        first, second = (c_real, c_fake) if d_loss else (c_fake, c_real)
        if cfg.avg:     # loss.py:90-94 / 104-108: 0.5 * (cal(first - mean(second), real) + cal(second - mean(first), fake))
            z1, z2 = ops.rel_logits(first, second, True), ops.rel_logits(second, first, True)
            g1, g2 = torch.empty_like(z1), torch.empty_like(z2)
            ops.fused_loss([dict(kind=kind, a=z1, grad=g1, target=real, weight=0.5 * lam, loss_weight=0.5, slot=slot),
                            dict(kind=kind, a=z2, grad=g2, target=fake, weight=0.5 * lam, loss_weight=0.5, slot=slot)],
                           self.losses)
            # d/d first = g1 - mean_batch(g2),  d/d second = g2 - mean_batch(g1): the same kernel in its forward form
            d_first, d_second = ops.rel_logits(g1, g2, True), ops.rel_logits(g2, g1, True)
        else:           # loss.py:96 / 110: cal(first - second, real)
            z = ops.rel_logits(first, second, False)
            d_first = torch.empty_like(z)
            ops.fused_loss([dict(kind=kind, a=z, grad=d_first, target=real, weight=lam, loss_weight=1.0, slot=slot)],
                           self.losses)
            d_second = ops.rel_logits(None, d_first, False, backward=True)
        return (d_first, d_second) if d_loss else (d_second, d_first)

    # ------------------------------------------------------------------------------------------
    def _step_segments(self, x, m, y):
        """The train step as a generator: it yields gradient-exchange requests `(bucket names, blocking[, wait names])`
        (data parallelism only) at the points where a bucket has become final; the consumer launches the all-reduce on the
        stream that is CURRENT at the yield -- so a request raised inside a lane only orders that lane behind the
        collective.  Everything between two yields is pure kernel launches; with the collectives launched through c10d's
        stream events the whole generator can be consumed inside ONE CUDA-graph capture."""
        cfg, rt = self.cfg, self.rt
        kind, real, fake = self._labels()
        multi = self.world > 1
        newg = lambda t: torch.empty_like(t)
        d1_params, d2_params = list(self.nets["D1"].parameters()), list(self.nets["D2"].parameters())
        # ================= D phase (cgan.py:278-305) =================
        L = self.lanes
        rt["D1"].zero_grads(); rt["D2"].zero_grads()
        # every distinct input concatenation (cgan.py:281-289, 321-324) is packed once per step and shared.
        # SGAN: each of the four discriminator passes is an independent chain  forward -> its loss term -> backward  (the
        # terms of D_loss are separable, and the backward passes only meet in atomic accumulations), so every pass runs on
        # its own lane and starts the moment its inputs exist: the two `real` passes at once, D1's `fake` pass after G1's
        # forward, D2's after G2's -- all of it underneath the generator chain on the main stream.  The only ordering kept
        # between the two passes of one discriminator is forward-after-forward: the BatchNorm running statistics are
        # order-dependent (real first, as in the reference).  Relativistic forms couple real and fake logits: there the
        # `fake` lane runs  fake forward -> objective -> both backward passes.
        self.losses.zero_()
        L.fork()

        def d_real(net, sources, packed, lam, slot):
            c, w = rt[net].forward(sources, True, packed=packed)
            after_fwd = L.record({"D1": 0, "D2": 1}[net])
            if cfg.rel:
                return c, w, after_fwd
            d = newg(c)
            ops.fused_loss([dict(kind=kind, a=c, grad=d, target=real, weight=0.5 * lam, loss_weight=0.5, slot=slot)],
                           self.losses)
            rt[net].backward(w, d, False)
            return c, None, after_fwd

        def d_fake(net, sources, packed, lam, slot, c_real, w_real):
            c, w = rt[net].forward(sources, True, packed=packed)
            if cfg.rel:
                d_r, d_f = self._rel_terms(c_real, c, True, lam, slot)
                rt[net].backward(w_real, d_r, False)
                rt[net].backward(w, d_f, False)
                return c
            d = newg(c)
            ops.fused_loss([dict(kind=kind, a=c, grad=d, target=fake, weight=0.5 * lam, loss_weight=0.5, slot=slot)],
                           self.losses)
            rt[net].backward(w, d, False)
            return c

        with L.lane(0):
            pk_xm = rt["D1"].pack_sources([x, m])
            c1r, w1r, ev1 = d_real("D1", [x, m], pk_xm, cfg.lambda2, 0)
        with L.lane(1):
            pk_xmy = rt["D2"].pack_sources([x, m, y])
            c2r, w2r, ev2 = d_real("D2", [x, m, y], pk_xmy, cfg.lambda3, 1)
        with self._critical():
            mp, wg1 = rt["G1"].forward([x], True)
            pk_xmp = rt["D1"].pack_sources([x, mp])
            L.lane_wait(2)
            L.wait_event(2, ev1)
            with L.lane(2):
                if cfg.rel:
                    L.wait_lane(2, 0)         # the objective needs C1_real; its backward pass runs here as well
                d_fake("D1", [x, mp], pk_xmp, cfg.lambda2, 0, c1r, w1r)
                # D1 is complete long before D2 (whose fake pass needs G2's forward): its gradient bucket goes on the wire
                # now, its share of optim_D.step (cgan.py:305) and its two G-phase forward passes (cgan.py:321-322, with
                # the UPDATED discriminator) follow on this lane, all underneath G2's forward and D2's fake chain
                ev_tick, c1r_g, c1f, w1f = None, None, None, None
                if self._early_d1:
                    L.wait_lane(2, 0)
                    if multi:
                        yield ("D1",), True
                    self.optim_D.step_partial(d1_params, tick=True, last=False)
                    ev_tick = L.record(2)
                    if cfg.rel or not cfg.skip_dead_real_passes:
                        c1r_g, _ = rt["D1"].forward([x, m], True, packed=pk_xm)       # cgan.py:321
                    c1f, w1f = rt["D1"].forward([x, mp], True, packed=pk_xmp)         # cgan.py:322
            share = rt["G2"].convs[0].thin == "cin" and rt["D1"].convs[0].thin == "cin"
            yp, wg2 = rt["G2"].forward([x, mp], True, packed=pk_xmp if share else None)
            pk_xmpyp = rt["D2"].pack_sources([x, mp, yp])
            L.lane_wait(3)
            L.wait_event(3, ev2)
            with L.lane(3):
                if cfg.rel:
                    L.wait_lane(3, 1)
                d_fake("D2", [x, mp, yp], pk_xmpyp, cfg.lambda3, 1, c2r, w2r)
        # D2's bucket is the one on the critical path: D2 fake chain -> all-reduce -> its Adam share -> G-phase D2 forwards
        self.last = dict(m_pred=mp, y_pred=yp)
        del w1r, w2r
        if self._early_d1:
            L.join_lanes((1, 3))
            if multi:
                yield ("D2",), True
            if ev_tick is not None:
                torch.cuda.current_stream().wait_event(ev_tick)      # the step counter advanced with D1's share
            self.optim_D.step_partial(d2_params, tick=False, last=True)           # cgan.py:305
        else:
            L.join()
            if multi:
                yield ("D1", "D2"), True
            self.optim_D.step()                                                   # cgan.py:305
        # ================= G phase (cgan.py:316-351) =================
        rt["G1"].zero_grads(); rt["G2"].zero_grads()
        # D2's two forward passes with the updated discriminator (cgan.py:323-324) sit on the step's critical path.  They are
        # independent except for the ORDER of their BatchNorm running-statistics updates (real first), so the real pass runs
        # on a lane next to the fake pass, whose updates are deferred and applied after the join (same arithmetic, same order)
        c2r_g, overlap_real = None, self._overlap_real
        need_real = cfg.rel or not cfg.skip_dead_real_passes
        # the thin layers' packed weights are re-derived from the updated parameters HERE, on the stream every consumer is
        # ordered behind -- not lazily inside whichever forward pass happens to run first (found with tools/noise_check.py: with
        # the real pass on a lane the fake pass could read D2's first / last layer weights while they were being re-packed)
        rt["D2"].ensure_packed()
        if not self._early_d1:
            L.fork()
            with L.lane(0):
                if need_real:
                    c1r_g, _ = rt["D1"].forward([x, m], True, packed=pk_xm)       # cgan.py:321
                c1f, w1f = rt["D1"].forward([x, mp], True, packed=pk_xmp)         # cgan.py:322
        if need_real and overlap_real:
            L.lane_wait(1)
            with L.lane(1):
                c2r_g, _ = rt["D2"].forward([x, m, y], True, packed=pk_xmy)      # cgan.py:323 (SGAN: running stats only)
            c2f, w2f = rt["D2"].forward([x, mp, yp], True, packed=pk_xmpyp, defer_running=True)   # cgan.py:324
            L.join_lanes((1,))
            rt["D2"].apply_deferred_running(w2f)
        else:
            if need_real:
                c2r_g, _ = rt["D2"].forward([x, m, y], True, packed=pk_xmy)
            c2f, w2f = rt["D2"].forward([x, mp, yp], True, packed=pk_xmpyp)
        L.join()                                                                  # D1's G-phase passes (lane 2)
        dm, dy = newg(mp), newg(yp)
        terms = [dict(kind=ops.KIND_L1, a=mp, b=m, grad=dm, weight=1.0, loss_weight=1.0, slot=4),
                 dict(kind=ops.KIND_L1, a=yp, b=y, grad=dy, weight=cfg.lambda1, loss_weight=1.0, slot=5)]
        if cfg.rel:
            ops.fused_loss(terms, self.losses)
            _, d1f = self._rel_terms(c1r_g, c1f, False, cfg.lambda2, 2)          # loss.py:102-110
            _, d2f = self._rel_terms(c2r_g, c2f, False, cfg.lambda3, 3)
        else:
            d1f, d2f = newg(c1f), newg(c2f)
            ops.fused_loss(terms + [
                dict(kind=kind, a=c1f, grad=d1f, target=real, weight=cfg.lambda2, loss_weight=1.0, slot=2),
                dict(kind=kind, a=c2f, grad=d2f, target=real, weight=cfg.lambda3, loss_weight=1.0, slot=3),
            ], self.losses)
        if self.visual_loss is not None and (cfg.lambda4 != 0 or cfg.lambda5 != 0):
            self._visual_terms(mp, yp, m, y, dm, dy)
        L.fork()
        with L.lane(0):
            di1 = rt["D1"].backward(w1f, d1f, True, param_grads=False)  # dgrad only: D is frozen (cgan.py:317-318)
        di2 = rt["D2"].backward(w2f, d2f, True, param_grads=False)
        ops.unpack_input_grad(di2, 4, 3, dy, True)
        ops.unpack_input_grad(di2, 3, 1, dm, True)
        L.join()
        ops.unpack_input_grad(di1, 3, 1, dm, True)
        with self._critical():
            dig2 = rt["G2"].backward(wg2, dy, True)
            ops.unpack_input_grad(dig2, 3, 1, dm, True)       # G2's input gradient, mask channel (cgan.py:286)
        if multi:
            yield ("G2",), False                              # async: overlaps G1's backward
        # G1's backward in two halves.  optim_G.step (cgan.py:351; an HBM-bound stream over 28 B/parameter) is applied per
        # gradient bucket on a lane underneath it: G2's share as soon as G2's (summed) gradients are there, i.e. under the
        # decoder half; after the decoder half G1's up-conv weight gradients (64 % of its parameters) are final, go on the
        # wire and are applied under the encoder half; only the rest of G1 (19.5 M parameters) is updated after the backward
        split = bool(L.streams)
        cap = int(os.environ.get("STCGAN_ADAM_OVERLAP_CTAS", "148"))
        g2_params = list(self.nets["G2"].parameters())
        early = split and self._adam_early
        if early:
            L.fork()
            with L.lane(0):
                if multi:
                    yield (), False, ("G2",)                  # this lane (not the backward chain) waits for G2's sum
                self.optim_G.step_partial(g2_params, tick=True, last=False, max_ctas=cap)
        with self._critical():
            rt["G1"].backward(wg1, dm, False, part="dec")
        if multi:
            yield ("G1.ups",), False
        if early:
            L.lane_wait(0)                                    # the lane sees the decoder half's weight gradients
            with L.lane(0):
                if multi:
                    yield (), False, ("G1.ups",)
                self.optim_G.step_partial(self._g1_ups, tick=False, last=False, max_ctas=cap)
        elif split:                                           # round-1 schedule: G2's share under the encoder half only
            L.fork()
            with L.lane(0):
                if multi:
                    yield (), False, ("G2",)
                self.optim_G.step_partial(g2_params, tick=True, last=False, max_ctas=cap)
        # the encoder half completes G1's weight gradients innermost level first: the four 512 x 512 down convs (86 % of what
        # is left) form the "deep" bucket, which goes on the wire -- and then through Adam, on the lane -- while the outer
        # levels are still running; only the small "tail" bucket is exchanged and applied after the backward pass
        deep = {"ev": None}

        def deep_done():                                      # (current stream = the weight-gradient stream)
            if multi:
                self._reduce((("G1.deep",), False), self._pending)
            deep["ev"] = torch.cuda.Event()
            deep["ev"].record()

        deep_early = early or (multi and split and self._deep_bucket)
        with self._critical():
            rt["G1"].backward(wg1, dm, False, part="enc", on_deep=deep_done if deep_early else None)
        if early:
            with L.lane(0):
                if deep["ev"] is not None:
                    L.streams[0].wait_event(deep["ev"])
                if multi:
                    yield (), False, ("G1.deep",)
                self.optim_G.step_partial(self._g1_deep, tick=False, last=False, max_ctas=cap)
        if split:
            L.join()
            if early:
                if multi:
                    yield ("G1.tail",), True
                self.optim_G.step_partial(self._g1_tail, tick=False, last=True)
            elif deep_early:                                  # (N > 1) only the small tail bucket is still to be exchanged
                yield ("G1.tail",), False, ("G1.ups",)
                self.optim_G.step_partial(self._g1_ups, tick=False, last=False)   # under the tail's all-reduce
                yield (), True
                self.optim_G.step_partial(self._g1_rest, tick=False, last=True)
            else:
                if multi:
                    yield ("G1.rest",), False, ("G1.ups",)
                self.optim_G.step_partial(self._g1_ups, tick=False, last=False)   # under the last bucket's all-reduce
                if multi:
                    yield (), True
                self.optim_G.step_partial(self._g1_rest, tick=False, last=True)
        else:
            if multi:
                yield ("G1.ups", "G1.rest"), True            # (blocking: waits for G2's sum as well)
            self.optim_G.step()                               # cgan.py:351
        for r in rt.values():
            r.ensure_packed()                                 # re-pack the updated weights for the next step

    def _visual_terms(self, mp, yp, m, y, dm, dy):
        """vis1 / vis2 of src/cgan.py:334-348 through the user-supplied torch callable: values into slots 6 / 7, gradients
        (torch autograd) added to dm / dy.  This is the one place where the engine runs foreign torch code."""
        cfg = self.cfg
        with torch.enable_grad():
            mp_, yp_ = mp.detach().requires_grad_(True), yp.detach().requires_grad_(True)
            v1 = self.visual_loss(mp_.expand(-1, 3, -1, -1), m.expand(-1, 3, -1, -1))
            v2 = self.visual_loss(yp_, y)
            g_m, g_y = torch.autograd.grad(cfg.lambda4 * v1 + cfg.lambda5 * v2, (mp_, yp_), allow_unused=True)
        self.losses[6:7].copy_(v1.detach().reshape(1))
        self.losses[7:8].copy_(v2.detach().reshape(1))
        if g_m is not None:
            dm.add_(g_m)
        if g_y is not None:
            dy.add_(g_y)

    def _reduce(self, req, pending):
        names, blocking = req[0], req[1]
        self.sync.reduce(names, blocking, pending, wait=req[2] if len(req) > 2 else ())

    # ------------------------------------------------------------------------------------------
    def train_step(self, x, m, y):
        """x [B,3,H,W], m [B,1,H,W], y [B,3,H,W]: float32 CUDA tensors in [-1,1].  Returns the device tensor of
        losses (index with SLOTS); nothing is synchronised with the host."""
        for t in (x, m, y):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError("train_step expects float32 CUDA tensors")
        self._pending = []
        for req in self._step_segments(x.contiguous(), m.contiguous(), y.contiguous()):
            self._reduce(req, self._pending)
        return self.losses

    def capture(self, x, m, y, warmup=3):
        """Capture the train step for inputs of this shape into ONE CUDA graph.  Under data parallelism the gradient
        all-reduces are part of the graph: NCCL's stream joins the capture through the events c10d records around every
        collective, so there are no host-side segment boundaries and the collectives overlap whatever else the graph has in
        flight (measured on 2 x B200 in round 1: 7.33 -> 6.69 ms per step against graph segments with host-driven
        all-reduces between them).  Such a graph holds NCCL work: call release_graphs() before destroy_process_group()
        (an atexit hook does it for engines that are still alive).  STCGAN_NCCL_IN_GRAPH=0 keeps the multi-GPU step eager
        (no graph at all; `replay` then simply runs train_step on the static inputs)."""
        self._static = tuple(t.contiguous().clone() for t in (x, m, y))
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.train_step(*self._static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.optim_D.prepare(); self.optim_G.prepare()
        before = _lib.launch_count()
        self._graphs, self._pending = [], []
        if self.world > 1 and os.environ.get("STCGAN_NCCL_IN_GRAPH", "1") == "0":
            self.train_step(*self._static)
            self.graph_launches = _lib.launch_count() - before
            self._graph = "eager"
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):            # the warm-up stream: per-stream scratch already exists
                for req in self._step_segments(*self._static):
                    self._reduce(req, self._pending)
            self._graphs.append(g)
            self.graph_launches = _lib.launch_count() - before
            self._graph = "graph"
        self._capture_state = self._state_signature()
        return self._graphs

    def _state_signature(self):
        return (self.optim_D.generation, self.optim_G.generation, tuple(id(n._rt) for n in self.nets.values()),
                sum(p._version for n in self.nets.values() for p in n.parameters()))

    def _validate_capture(self):
        """A captured graph bakes in device addresses (weights, packed bf16 copies, optimiser tables and state).  Before a
        replay: (i) optimiser tables that were rebuilt, or modules that were moved / cast, invalidate the graph -> loud
        error; (ii) weights that were overwritten in place since (load_state_dict of a checkpoint, src/cgan.py:511-542) only
        leave the packed bf16 copies stale -> they are re-packed here, into the same buffers."""
        sig = self._state_signature()
        if sig == self._capture_state:
            return
        if sig[:3] != self._capture_state[:3]:
            raise RuntimeError("the optimiser tables or the module runtimes changed since capture() (set_packed_grads / "
                               "load of a state with a different layout / .to()): call capture() again")
        for r in self.rt.values():
            r.ensure_packed()
        self._capture_state = self._state_signature()

    def release_graphs(self):
        """Drop the captured graphs (and the NCCL work they may hold) -- call before destroy_process_group()."""
        if self._graphs:
            torch.cuda.synchronize(self.device)
        self._graphs, self._graph = [], None

    def replay(self, x=None, m=None, y=None):
        """Run the captured step (optionally on new inputs of the captured shape)."""
        if not self._graph:
            raise RuntimeError("call capture() first")
        # lr / grad_scale live in a device vector the captured Adam kernels read through a pointer: push scheduler changes
        # (ExponentialLR once per epoch, src/cgan.py:383-384) before the launch
        self.optim_D.sync_hyper(); self.optim_G.sync_hyper()
        self._validate_capture()
        for dst, src in zip(self._static, (x, m, y)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        if self._graph == "eager":
            return self.train_step(*self._static)
        for g in self._graphs:
            g.replay()
        self.optim_D.bump_host_counters(); self.optim_G.bump_host_counters()
        return self.losses

    def replay_async(self, x, m, y):
        """Pipelined form of `replay` for a training loop that feeds host batches: the pinned host tensors x, m, y are copied
        into a double-buffered device inbox on a copy stream (the transfer of batch i+1 runs underneath the compute of
        batch i), the captured step first moves its inbox into the graph's static inputs, and the step's losses are copied
        to pinned host memory behind it.  Returns the losses of the PREVIOUS call as a host tensor (None on the first call)
        -- the only host synchronisation is on a step that has had a whole step's time to finish; `flush()` returns the
        last one."""
        if not self._graph:
            raise RuntimeError("call capture() first")
        if getattr(self, "_pipe", None) is None:
            mk = lambda: tuple(torch.empty_like(t) for t in self._static)
            self._pipe = dict(i=0, inbox=[mk(), mk()], copy=torch.cuda.Stream(device=self.device),
                              loss=[torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)],
                              ready=[torch.cuda.Event() for _ in range(2)], consumed=[torch.cuda.Event() for _ in range(2)],
                              done=[torch.cuda.Event() for _ in range(2)], primed=[False, False])
        P = self._pipe
        i = P["i"]
        main = torch.cuda.current_stream()
        if P["primed"][i]:
            P["copy"].wait_event(P["consumed"][i])        # the step that last used this inbox slot has copied it out
        with torch.cuda.stream(P["copy"]):
            for dst, src in zip(P["inbox"][i], (x, m, y)):
                dst.copy_(src, non_blocking=True)
            P["ready"][i].record(P["copy"])
        main.wait_event(P["ready"][i])
        for dst, src in zip(self._static, P["inbox"][i]):
            dst.copy_(src, non_blocking=True)
        P["consumed"][i].record(main)
        self.replay()
        P["loss"][i].copy_(self.losses, non_blocking=True)
        P["done"][i].record(main)
        P["primed"][i] = True
        P["i"] = 1 - i
        j = 1 - i
        if not P["primed"][j]:
            return None
        P["done"][j].synchronize()
        return P["loss"][j].clone()

    def flush(self):
        """Host copy of the losses of the last `replay_async` call (waits for that step)."""
        P = getattr(self, "_pipe", None)
        if P is None:
            return None
        j = 1 - P["i"]
        if not P["primed"][j]:
            return None
        P["done"][j].synchronize()
        return P["loss"][j].clone()

    def replay_u8(self, x8, m8, y8):
        """Captured step fed with decoded uint8 HWC images ([B,H,W,3], [B,H,W,1], [B,H,W,3]; host-pinned or device): the
        dataset's uint8 -> [-1,1] float CHW transform (src/dataset.py:100-110,152) runs on the GPU."""
        if not self._graph:
            raise RuntimeError("call capture() first")
        if getattr(self, "_static_u8", None) is None:
            self._static_u8 = tuple(torch.empty(t.shape, dtype=torch.uint8, device=self.device) for t in (x8, m8, y8))
        for dev8, src, dst in zip(self._static_u8, (x8, m8, y8), self._static):
            dev8.copy_(src, non_blocking=True)
            ops.u8_to_nchw(dev8, out=dst)
        return self.replay()

    def loss_dict(self, losses=None):
        v = (self.losses if losses is None else losses).detach().cpu().tolist()
        d = dict(zip(SLOTS, v))
        c = self.cfg
        d["D_loss"] = c.lambda2 * d["D1_loss"] + c.lambda3 * d["D2_loss"]
        d["G_loss"] = (d["data1_loss"] + c.lambda1 * d["data2_loss"] + c.lambda2 * d["G1_loss"] + c.lambda3 * d["G2_loss"]
                       + c.lambda4 * d["vis1_loss"] + c.lambda5 * d["vis2_loss"])
        return d


@torch.no_grad()
def infer(G1, G2, x, quantize=True, want_float=True):
    """CGAN.infer core (src/cgan.py:437-446 + utils.float2uint): eval-mode G1 -> G2; returns
    (m_pred, y_pred [NCHW fp32], m_u8, y_u8 [N,H,W,C] uint8 or None).  On the bf16 path the uint8 images come out of the
    last layers' epilogues; `want_float=False` (uint8-only callers) then skips y_pred's float output (y_pred is None;
    m_pred is always produced, it is G2's input)."""
    if G1.training or G2.training:
        raise RuntimeError("infer() expects G1.eval(); G2.eval() like the reference (cgan.py:422-423)")
    rt1, rt2 = G1.runtime(), G2.runtime()
    x = x.contiguous()
    if os.environ.get("STCGAN_INFER_FOLD", "1") != "0":
        if quantize:                                # ... and the uint8 quantisation into the last layers' epilogues
            mp, m8 = rt1.forward_inference([x], quantize=True)
            yp, y8 = rt2.forward_inference([x, mp], quantize=True, want_float=want_float)
            return mp, yp, m8, y8
        mp = rt1.forward_inference([x])             # BatchNorm(eval) + activations folded into the conv epilogues
        yp = rt2.forward_inference([x, mp])
    else:
        mp, _ = rt1.forward([x], False)
        yp, _ = rt2.forward([x, mp], False)
    if not quantize:
        return mp, yp, None, None
    return mp, yp, ops.float2uint_hwc(mp), ops.float2uint_hwc(yp)


@torch.no_grad()
def infer_u8(G1, G2, x_u8, out_m=None, out_y=None):
    """CGAN.infer end to end on decoded images: `x_u8` = uint8 [N,H,W,3] as cv2 gives it (host-pinned or device).  The
    dataset transform (src/dataset.py:100-110,152), G1 -> G2 (cgan.py:437-438) and utils.float2uint (cgan.py:441-446) all run
    on the GPU; returns (m_u8 [N,H,W,1], y_u8 [N,H,W,3]) on the device, or -- when pinned host buffers `out_m` / `out_y` are
    given -- copies them there asynchronously (synchronise the stream before reading them)."""
    dev = next(G1.parameters()).device
    x8 = x_u8 if x_u8.is_cuda else x_u8.to(dev, non_blocking=True)
    x = ops.u8_to_nchw(x8)
    _, _, m8, y8 = infer(G1, G2, x, want_float=False)
    if out_m is not None:
        out_m.copy_(m8, non_blocking=True)
    if out_y is not None:
        out_y.copy_(y8, non_blocking=True)
    return m8, y8


class InferencePipeline:
    """CGAN.infer over a stream of host batches (src/cgan.py:426-460) with the transfers off the critical path: while batch i
    runs G1 -> G2 on the compute stream, batch i+1's decoded uint8 images travel host -> device and batch i-1's uint8 results
    travel device -> host on two copy streams (double-buffered pinned / device slots).  `submit(x_u8)` returns the PREVIOUS
    batch's `(m_u8 [N,H,W,1], y_u8 [N,H,W,3])` as pinned host tensors (None on the first call), `flush()` the last one's;
    a returned pair stays valid until the second next `submit`."""

    def __init__(self, G1, G2):
        self.G1, self.G2 = G1, G2
        self.dev = next(G1.parameters()).device
        self.cin, self.cout = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)
        self.slots = [dict(primed=False) for _ in range(2)]
        self.k = 0

    def _slot(self, i, x_u8):
        sl = self.slots[i]
        if sl.get("shape") != tuple(x_u8.shape):
            n, h, w, _ = x_u8.shape
            sl.update(shape=tuple(x_u8.shape), x8=torch.empty(x_u8.shape, dtype=torch.uint8, device=self.dev),
                      out_m=torch.empty((n, h, w, 1), dtype=torch.uint8).pin_memory(),
                      out_y=torch.empty((n, h, w, self.G2.out_channels), dtype=torch.uint8).pin_memory(),
                      ready=torch.cuda.Event(), consumed=torch.cuda.Event(), computed=torch.cuda.Event(), done=torch.cuda.Event(),
                      primed=False, keep=None)
        return sl

    @torch.no_grad()
    def submit(self, x_u8):
        if x_u8.is_cuda or x_u8.dtype != torch.uint8 or x_u8.dim() != 4:
            raise ValueError("submit expects a (pinned) host uint8 [N,H,W,3] batch")
        i = self.k % 2
        sl, main = self._slot(i, x_u8), torch.cuda.current_stream(self.dev)
        if sl["primed"]:
            sl["done"].synchronize()                   # this slot's previous results have left the device (host may reuse out_*)
            self.cin.wait_event(sl["consumed"])
        with torch.cuda.stream(self.cin):
            sl["x8"].copy_(x_u8, non_blocking=True)
            sl["ready"].record(self.cin)
        main.wait_event(sl["ready"])
        x = ops.u8_to_nchw(sl["x8"])
        sl["consumed"].record(main)
        _, _, m8, y8 = infer(self.G1, self.G2, x, want_float=False)
        sl["computed"].record(main)
        self.cout.wait_event(sl["computed"])
        with torch.cuda.stream(self.cout):
            sl["out_m"].copy_(m8, non_blocking=True)
            sl["out_y"].copy_(y8, non_blocking=True)
            sl["done"].record(self.cout)
        m8.record_stream(self.cout); y8.record_stream(self.cout)
        sl["primed"] = True
        self.k += 1
        return self._result((self.k - 2) % 2) if self.k >= 2 else None

    def _result(self, j):
        sl = self.slots[j]
        if not sl.get("primed"):
            return None
        sl["done"].synchronize()
        return sl["out_m"], sl["out_y"]

    def flush(self):
        return self._result((self.k - 1) % 2) if self.k >= 1 else None
