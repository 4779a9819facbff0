"""Fused multi-tensor Adam with torch.optim.Adam's state layout.

Replaces `optim.Adam(params, lr, betas=(beta1, beta2))` of src/cgan.py:85-90 (eps 1e-8, no weight decay,
no amsgrad).  `state_dict()` / `load_state_dict()` are inherited from torch.optim.Optimizer and produce the
same structure as torch's Adam (`state[i] = {step, exp_avg, exp_avg_sq}`), so `checkpoint.tar`
(src/cgan.py:490-523) round-trips.  One kernel launch updates every tensor of a param group; the step
counter and bias corrections live on the device so the whole train step can be replayed as a CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam implements the configuration the reference uses (no weight decay / amsgrad)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._tables = {}        # group index -> dict(sig, table, blocks, nblocks, keep, hyper, hyper_host)
        self._packed_grads = {}  # id(param) -> (flat fp32 view, d0, d1): engine-provided packed gradients
        self._pack_targets = {}  # id(param) -> ConvOp whose bf16 packed copies the Adam kernel refreshes in the same pass
        self.grad_scale = 1.0
        self._order_key = None   # optional: parameter -> sort key of its blocks in the device table (see set_block_order)
        # Every (re)build of a device table bumps `generation`.  A captured CUDA graph bakes in the addresses of the table,
        # the block list, the hyper-parameter vector and the exp_avg / exp_avg_sq tensors, so whoever captured `step()`
        # records the generation and must not replay the graph after it changed (STCGANEngine.replay checks this).
        self.generation = 0

    def set_packed_grads(self, views):
        """Engine hook: gradients that live in packed [16][d0][d1] layout instead of `p.grad`."""
        self._packed_grads = dict(views)
        self._tables.clear()

    def set_block_order(self, key):
        """Engine hook: order the tensors of the device table by `key(param)` (stable), so that groups of parameters that
        are updated by separate `step_partial` calls (one per gradient bucket) occupy contiguous block ranges."""
        self._order_key = key
        self._tables.clear()

    def set_pack_targets(self, convs):
        """Engine hook: ConvOps (with .weight, .p1, .p2, .mark_packed()) whose packed bf16 weights are rewritten by the
        optimiser kernel itself, so no separate pack pass runs after the step."""
        self._pack_targets = {id(c.weight): c for c in convs}
        self._tables.clear()

    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict, but state that already lives on the device is updated IN PLACE: the
        exp_avg / exp_avg_sq tensors (and with them the device tables that point at them) keep their addresses, so a
        checkpoint (src/cgan.py:511-523) can be restored after the train step was captured into a CUDA graph.  The device
        step counter / bias corrections are refreshed from the loaded `step`."""
        old = {p: dict(st) for p, st in self.state.items() if "exp_avg" in st}
        super().load_state_dict(state_dict)
        kept = True
        for p, st_old in old.items():
            st = self.state.get(p)
            if st is None or "exp_avg" not in st:
                kept = False
                continue
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].shape == st_old[k].shape and st_old[k].dtype == torch.float32:
                    st_old[k].copy_(st[k])
                    st[k] = st_old[k]
                else:
                    kept = False
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32)
        if not kept or not old:
            self._tables.clear()
            return
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is None:
                continue
            if t["sig"] != self._signature(group):
                self._tables.pop(gi)
                continue
            t["keep"] = [(p, g, self.state[p]) for p, g, _ in t["keep"]]      # the state dicts are new objects
            b1, b2 = group["betas"]
            steps_done = float(t["keep"][0][2]["step"])
            t["hyper_host"] = [float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(self.grad_scale),
                               steps_done, 0.0, 0.0]
            t["hyper"].copy_(torch.tensor(t["hyper_host"], dtype=torch.float32))

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)     # host counter, like torch's default Adam
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _grad_of(self, p):
        if id(p) in self._packed_grads:
            return self._packed_grads[id(p)]
        if p.grad is not None:
            return p.grad, 0, 0
        return None

    def _signature(self, group):
        sig = []
        for p in group["params"]:
            g = self._grad_of(p)
            if g is None:
                continue
            st = self.state.get(p, {})
            if "exp_avg" not in st:
                return None
            sig.append((p.data_ptr(), g[0].data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
        return tuple(sig)

    def _build_table(self, group):
        chunk = _lib.load().stcgan_adam_chunk()
        tile = _lib.load().stcgan_adam_tile()
        entries, blocks, keep, fused, ranges = [], [], [], [], {}
        params = list(group["params"])
        if self._order_key is not None:
            params.sort(key=self._order_key)
        for p in params:
            gd = self._grad_of(p)
            if gd is None:
                continue
            g, d0, d1 = gd
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()):
                raise RuntimeError("FusedAdam handles contiguous float32 CUDA parameters/gradients only (no CPU path)")
            st = self._state_for(p)
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].device != p.device or not st[k].is_contiguous():
                    st[k] = st[k].to(p.device).contiguous()
            ti = len(entries)
            conv = self._pack_targets.get(id(p))
            p1 = p2 = None
            aligned = all(t.data_ptr() % 16 == 0 for t in (p, g, st["exp_avg"], st["exp_avg_sq"]))
            if (conv is not None and d0 > 0 and d0 % tile == 0 and d1 % tile == 0 and conv.p1 is not None
                    and conv.p1.dtype == torch.bfloat16 and aligned):
                p1, p2 = conv.p1.data_ptr(), conv.p2.data_ptr()
                fused.append(conv)
            entries.append(_lib.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                           st["exp_avg_sq"].data_ptr(), p.numel(), d0, d1, p1, p2))
            keep.append((p, g, st))
            nblk = (d0 // tile) * (d1 // tile) if p1 is not None else (p.numel() + chunk - 1) // chunk
            ranges[id(p)] = (len(blocks), nblk)
            blocks += [(ti, c) for c in range(nblk)]
        if not entries:
            return None
        arr = (_lib.AdamTensor * len(entries))(*entries)
        dev = keep[0][0].device
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        blk = torch.from_numpy(np.asarray(blocks, dtype=np.int32).reshape(-1).copy()).to(dev)
        b1, b2 = group["betas"]
        steps_done = float(keep[0][2]["step"].item())
        hyper_host = [float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(self.grad_scale), steps_done, 0.0, 0.0]
        hyper = torch.tensor(hyper_host, dtype=torch.float32, device=dev)
        return dict(sig=self._signature(group), table=table, blocks=blk, nblocks=len(blocks), keep=keep,
                    hyper=hyper, hyper_host=hyper_host, fused=fused, ranges=ranges)

    def _sync_hyper(self, group, t):
        """Push lr / grad_scale changes (ExponentialLR steps once per epoch, src/cgan.py:383-384) to the device."""
        want = [float(group["lr"]), t["hyper_host"][1], t["hyper_host"][2], t["hyper_host"][3], float(self.grad_scale)]
        if want != t["hyper_host"][:5]:
            t["hyper_host"][:5] = want
            t["hyper"][:5].copy_(torch.tensor(want, dtype=torch.float32), non_blocking=False)

    def prepare(self):
        """Build the device tables now (e.g. before CUDA-graph capture)."""
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is None or t["sig"] != self._signature(group):
                self._tables[gi] = self._build_table(group)
                self.generation += 1

    def sync_hyper(self):
        """Push lr / grad_scale of every param group to the device vector the kernels read (a plain H2D copy outside any
        graph: the vector is read through its pointer, so a captured `step()` picks the new values up on its next replay).
        STCGANEngine.replay calls this before every replay -- the per-epoch ExponentialLR of src/cgan.py:383-384 then acts on
        the captured path exactly as on the eager one."""
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is not None:
                self._sync_hyper(group, t)

    def bump_host_counters(self):
        """Account on the host for one device-side step (used after a CUDA-graph replay of `step`)."""
        for t in self._tables.values():
            if t is not None:
                for _, _, st in t["keep"]:
                    st["step"] += 1

    @torch.no_grad()
    def step_partial(self, params, tick, last, max_ctas=0):
        """Update only `params` (a contiguous run of this optimiser's single param group, e.g. one network's parameters).
        One optimiser step = several partial calls covering all parameters: `tick=True` on the first (advances the step
        counter on the device), `last=True` on the final one (host-side bookkeeping).  Calls after the first must be
        stream-ordered after it."""
        lib = _lib.load()
        self.prepare()
        if len(self.param_groups) != 1:
            raise RuntimeError("step_partial expects a single param group")
        group, t = self.param_groups[0], self._tables.get(0)
        if t is None:
            return
        self._sync_hyper(group, t)
        rs = [t["ranges"][id(p)] for p in params if id(p) in t["ranges"]]
        first = min(r[0] for r in rs)
        end = max(r[0] + r[1] for r in rs)
        if sum(r[1] for r in rs) != end - first:
            raise RuntimeError("step_partial: the parameters do not form a contiguous run of the param group")
        _lib.check(lib.stcgan_adam_step_range(t["table"].data_ptr(), t["blocks"].data_ptr(), first, end - first,
                                              t["hyper"].data_ptr(), int(tick), int(max_ctas),
                                              torch.cuda.current_stream().cuda_stream),
                   "stcgan_adam_step_range")
        # host bookkeeping for the tensors of THIS launch: version bump (whoever caches derived copies -- the thin layers'
        # packed weights -- re-derives them on next use) and "p1 / p2 are fresh" for the tensors the kernel re-packed itself
        ids = {id(p) for p in params}
        for p, _, _ in t["keep"]:
            if id(p) in ids:
                torch.autograd.graph.increment_version(p)
        for conv in t["fused"]:
            if id(conv.weight) in ids:
                conv.mark_packed()
        # (during graph capture no kernel runs: the host step counter must not advance -- it would run ahead of the device
        # counter hyper[5])
        if last and not torch.cuda.is_current_stream_capturing():
            for _, _, st in t["keep"]:
                st["step"] += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        self.prepare()
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is None:
                continue
            self._sync_hyper(group, t)
            _lib.check(lib.stcgan_adam_step(t["table"].data_ptr(), t["blocks"].data_ptr(), t["nblocks"],
                                            t["hyper"].data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "stcgan_adam_step")
            counting = not torch.cuda.is_current_stream_capturing()      # see step_partial
            for p, _, st in t["keep"]:
                if counting:
                    st["step"] += 1
                torch.autograd.graph.increment_version(p)
            for conv in t["fused"]:
                conv.mark_packed()          # the kernel above already rewrote conv.p1 / conv.p2
        return loss
