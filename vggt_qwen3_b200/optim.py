"""Fused AdamW for the projector parameter group.

The reference builds `torch.optim.AdamW([{base_params, lr}, {projector_params, proj_lr}], weight_decay=...)`
(src/train/train_sft.py:138-156) and steps it after every accumulation window (:232-236).  For the 1.23 B-parameter
Perceiver that step is 34 GB of HBM traffic; `FusedAdamW` does it in one pass per tensor (`vq_adamw_step`,
csrc/trainops.cu) and writes the bf16 compute copy of each matrix in the same pass, so the next forward does not re-cast
the weights.  Same constructor arguments, `param_groups`, `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`) and
LR-scheduler interplay as `torch.optim.AdamW`; CUDA f32 parameters only -- there is no CPU fallback.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_scale: float = 1.0) -> None:
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        if grad_scale == 0:
            raise ValueError("FusedAdamW: grad_scale must be non-zero")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale
        self._projector = None

    def attach_projector(self, projector) -> "FusedAdamW":
        """Let step() refresh `projector`'s bf16 compute copies in the same pass (PerceiverProjector only)."""
        self._projector = projector
        return self

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        copies = {}
        cc = None
        if self._projector is not None and self._projector._compute is not None:
            cc = self._projector._compute
            for i, p in enumerate(self._projector._ordered_params()):
                copies[id(p)] = i
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: contiguous CUDA f32 parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)   # like torch.optim.AdamW: a (CPU) tensor
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                bf16_copy: Optional[torch.Tensor] = None
                idx = copies.get(id(p))
                if idx is not None and cc["bufs"][idx].dtype == torch.bfloat16 and cc["device"] == p.device:
                    bf16_copy = cc["bufs"][idx]
                ops.adamw_step(p, st["exp_avg"], st["exp_avg_sq"], grad, int(st["step"]), lr=float(group["lr"]),
                               betas=group["betas"], eps=group["eps"], weight_decay=group["weight_decay"],
                               grad_scale=self.grad_scale, param_bf16=bf16_copy)
                if idx is not None:
                    # the kernel wrote through raw pointers (no autograd version bump): mark the compute copy current
                    # if it was refreshed in the same pass, stale otherwise (f32 vector copies are re-copied lazily)
                    cc["versions"][idx] = (p._version, p.data_ptr()) if bf16_copy is not None else -1
        return loss


class ShardedAdamW:
    """ZeRO-1 style data-parallel AdamW for a `PerceiverProjector`: per step

        reduce-scatter of the gradient  ->  AdamW on the shard this rank owns  ->  all-gather of the updated weights

    The reference's equivalent is DeepSpeed's reduce-scatter / all-gather around its optimizer
    (configs/deepspeed_zero3.json:7-14, src/train/train_sft.py:138-163,217).  Here:

      * the projector writes its gradient straight into this optimizer's wire buffer G -- matrices in bf16 from the
        fp32 TMEM accumulators of the wgrad GEMMs, vectors in f32 (`PerceiverProjector.use_flat_storage`): no f32
        gradient buffer, no cast pass, `p.grad` stays None;
      * G and the compute copies W (what the forward kernels read: bf16 matrices, f32 vectors) live in ONE
        NVLink-symmetric allocation per rank; rank r owns the r-th slice of both regions;
      * step(), on a side stream: barrier | the peers' slices of G are pulled over NVLink by the copy engines |
        ONE kernel per region (`vq_adamw_reduce_step`) sums the N slices in fp32, applies AdamW to the f32 master /
        exp_avg / exp_avg_sq slices (only 1/N of the optimizer state exists per rank) and writes the updated slice of
        W | barrier | the peers' updated slices of W are pulled | barrier;
      * the next Perceiver forward waits for that stream (`projector.before_forward`); everything that does not depend
        on the projector weights -- the frozen aggregator forward of the next step -- overlaps with it.

    The module's f32 `nn.Parameter`s are NOT touched by step() (they would cost a second 4.9 GB write per step): call
    `sync_parameters()` before reading or saving them.  `accumulate_steps > 1` keeps the projector on its f32
    gradient buffer (in-kernel accumulation) and casts it into G at the step boundary.
    world_size 1 degenerates to a fused single-launch-per-region AdamW on flat buffers."""

    def __init__(self, projector, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, *, group=None, accumulate_steps: int = 1, grad_scale: float = 1.0,
                 extra_params: Iterable = (), transport=None, kernels=None) -> None:
        import torch.distributed as dist

        from . import dist as vdist

        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("ShardedAdamW: invalid hyper-parameter")
        self.projector = projector
        self.param_groups = [dict(params=list(projector._ordered_params()), lr=lr, betas=betas, eps=eps,
                                  weight_decay=weight_decay)]   # LR schedulers read / write param_groups[0]["lr"]
        self.accumulate_steps = max(1, int(accumulate_steps))
        self.grad_scale = float(grad_scale)
        self._kernels = kernels or _CudaKernels()
        params = projector._ordered_params()
        dev = params[0].device
        lay = dict(projector.flat_layout())
        # small trainable tensors outside the projector (geom_head, src/train/train_sft.py:139-145 puts them in the same
        # parameter group): appended to the f32 vector region; their autograd .grad is copied in at step()
        self.extra = [p for p in extra_params]
        self._extra_off = []
        for p in self.extra:
            self._extra_off.append(lay["vec"])
            lay["vec"] += (p.numel() + 63) // 64 * 64
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        # region sizes padded so that every rank owns an equal, 128-element-aligned slice
        self.slice = {k: ((lay[k] + world - 1) // world + 127) // 128 * 128 for k in ("mat", "vec")}
        self.region = {k: self.slice[k] * world for k in ("mat", "vec")}
        g_bytes = self.region["mat"] * 2 + self.region["vec"] * 4
        if transport is None:
            transport = (vdist.SymmetricMemoryTransport(2 * g_bytes, dev, group) if world > 1
                         else vdist.LocalTransport(2 * g_bytes, dev))
        elif callable(transport):
            transport = transport(2 * g_bytes, dev, group)
        self.transport = transport
        self.world, self.rank = transport.world, transport.rank
        self._g_bytes = g_bytes
        G, W = self._carve(transport.local(), 0), self._carve(transport.local(), g_bytes)
        self.G, self.W = G, W
        r = self.rank
        self.master, self.exp_avg, self.exp_avg_sq, self.stage = {}, {}, {}, {}
        flat_params = {k: torch.zeros(self.region[k], dtype=torch.float32, device=dev) for k in ("mat", "vec")}
        for p, kind, off in zip(params, lay["kinds"], lay["offsets"]):
            flat_params[kind][off:off + p.numel()].copy_(p.detach().to(torch.float32).reshape(-1))
        for p, off in zip(self.extra, self._extra_off):
            flat_params["vec"][off:off + p.numel()].copy_(p.detach().to(torch.float32).reshape(-1))
        for k in ("mat", "vec"):
            sl = slice(r * self.slice[k], (r + 1) * self.slice[k])
            self.master[k] = flat_params[k][sl].clone()
            self.exp_avg[k] = torch.zeros_like(self.master[k])
            self.exp_avg_sq[k] = torch.zeros_like(self.master[k])
            gdt = torch.bfloat16 if k == "mat" else torch.float32
            self.stage[k] = torch.empty((max(self.world - 1, 1), self.slice[k]), dtype=gdt, device=dev)
        # compute copies start as the cast of the current parameters (all ranks hold identical parameters)
        W["mat"].copy_(flat_params["mat"])
        W["vec"].copy_(flat_params["vec"])
        del flat_params
        self._lay = lay
        self._param_versions = [(p._version, p.data_ptr()) for p in params]
        direct = self.accumulate_steps == 1
        projector.use_flat_storage(W["mat"], W["vec"], G["mat"] if direct else None, G["vec"] if direct else None)
        projector.before_forward = self.wait
        self._mark_compute_current()
        self.step_count = 0
        self._micro = 0
        self.peers = [(r + i) % self.world for i in range(1, self.world)]   # staggered: no hot spot
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._done = None
        self.bytes_pulled = 0

    # ------------------------------------------------------------------ layout helpers
    def _carve(self, buf: torch.Tensor, base: int) -> dict:
        nm, nv = self.region["mat"], self.region["vec"]
        return {"mat": buf[base:base + nm * 2].view(torch.bfloat16),
                "vec": buf[base + nm * 2:base + nm * 2 + nv * 4].view(torch.float32)}

    def _mark_compute_current(self) -> None:
        """The compute copies in W are authoritative: stop the projector from re-casting them from its (stale) f32
        nn.Parameters, except after a load_state_dict / manual edit (version bump), which reload_masters() absorbs."""
        pr = self.projector
        dev = self.W["mat"].device
        cc = pr._refresh_compute_copies(dev) if dev.type == "cuda" else None
        if cc is not None:
            for i, p in enumerate(pr._ordered_params()):
                cc["versions"][i] = (p._version, p.data_ptr())

    def _latents_to_bf16(self) -> None:
        """The latents' compute copy is bf16 but they live (f32) in the vector region: re-cast after every update."""
        pr = self.projector
        if pr._compute is None:
            return
        lat = pr._compute["bufs"][0]
        off = self._lay["offsets"][0]
        self._kernels.cast_f32_bf16(self.W["vec"][off:off + lat.numel()], lat)

    # ------------------------------------------------------------------ optimizer protocol
    def zero_grad(self, set_to_none: bool = True) -> None:
        self.projector.mark_grads_consumed()
        for p in self.extra:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
        if self.accumulate_steps > 1:
            self.projector.zero_grad(set_to_none=set_to_none)
        self._micro = 0

    def wait(self) -> None:
        """Make the current stream wait for the step that is still in flight on the optimizer's stream."""
        if self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)

    def reload_masters(self) -> None:
        """Re-read this rank's master slices (and the compute copies) from the module's f32 parameters, e.g. after
        `load_state_dict`; the optimizer moments are kept."""
        params = self.projector._ordered_params()
        for p, kind, off in zip(params, self._lay["kinds"], self._lay["offsets"]):
            src = p.detach().to(torch.float32).reshape(-1)
            self.W[kind][off:off + p.numel()].copy_(src)
            lo, hi = self.rank * self.slice[kind], (self.rank + 1) * self.slice[kind]
            a, b = max(off, lo), min(off + p.numel(), hi)
            if a < b:
                self.master[kind][a - lo:b - lo].copy_(src[a - off:b - off])
        self._param_versions = [(p._version, p.data_ptr()) for p in params]
        self._mark_compute_current()
        self._latents_to_bf16()

    @torch.no_grad()
    def step(self) -> None:
        """Call once per optimizer step (at the accumulation boundary).  Asynchronous: returns after enqueueing."""
        params = self.projector._ordered_params()
        if [(p._version, p.data_ptr()) for p in params] != self._param_versions:
            self.reload_masters()
        self.step_count += 1
        grp = self.param_groups[0]
        cur = torch.cuda.current_stream() if self.stream is not None else None
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.stream.wait_event(ev)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()
        with ctx:
            if self.accumulate_steps > 1:        # f32 accumulation buffer -> wire buffer (one cast + one copy)
                flat = self.projector.grad_flat
                if flat is None:
                    raise RuntimeError("ShardedAdamW.step(): no gradient has been computed")
                plain = self.projector.flat_layout()      # the projector's own regions (without the extra tensors)
                nm, nv = plain["mat"], plain["vec"]
                self._kernels.cast_f32_bf16(flat[:nm], self.G["mat"][:nm])
                self.G["vec"][:nv].copy_(flat[nm:nm + nv])
            for p, off in zip(self.extra, self._extra_off):      # geom_head & co: autograd gradients -> wire buffer
                dst = self.G["vec"][off:off + p.numel()]
                if p.grad is None:
                    dst.zero_()
                else:
                    dst.copy_(p.grad.detach().reshape(-1))
            r, N = self.rank, self.world
            if N > 1:
                self.transport.barrier()         # every rank's gradient is complete
                for i, p in enumerate(self.peers):
                    pg = self._carve(self.transport.peer(p), 0)
                    for k in ("mat", "vec"):
                        self.stage[k][i].copy_(pg[k][r * self.slice[k]:(r + 1) * self.slice[k]])
                        self.bytes_pulled += self.stage[k][i].numel() * self.stage[k][i].element_size()
            scale = self.grad_scale / (N * self.accumulate_steps)
            for k in ("mat", "vec"):
                sl = slice(r * self.slice[k], (r + 1) * self.slice[k])
                srcs = [self.G[k][sl]] + [self.stage[k][i] for i in range(N - 1)]
                self._kernels.adamw_reduce_step(self.master[k], self.exp_avg[k], self.exp_avg_sq[k], srcs, self.W[k][sl],
                                                self.step_count, lr=float(grp["lr"]), betas=grp["betas"],
                                                eps=grp["eps"], weight_decay=grp["weight_decay"], grad_scale=scale)
            if N > 1:
                self.transport.barrier()         # every slice of W is updated; nobody reads G any more
                for p in self.peers:
                    pw = self._carve(self.transport.peer(p), self._g_bytes)
                    for k in ("mat", "vec"):
                        sl = slice(p * self.slice[k], (p + 1) * self.slice[k])
                        self.W[k][sl].copy_(pw[k][sl])
                        self.bytes_pulled += self.slice[k] * self.W[k].element_size()
                self.transport.barrier()         # nobody reads this rank's W slice any more
            self._latents_to_bf16()
            for p, off in zip(self.extra, self._extra_off):      # the vector region is f32: exact write-back
                p.data.copy_(self.W["vec"][off:off + p.numel()].view(p.shape))
            if self.stream is not None:
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
        self.projector.mark_grads_consumed()
        self._micro = 0

    @torch.no_grad()
    def sync_parameters(self) -> None:
        """Write the current weights back into the module's f32 nn.Parameters (checkpointing).  Every rank gathers the
        f32 master slices (torch.distributed all_gather; rare, not on the step path)."""
        import torch.distributed as dist

        self.wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        full = {}
        for k in ("mat", "vec"):
            if self.world > 1:
                parts = [torch.empty_like(self.master[k]) for _ in range(self.world)]
                dist.all_gather(parts, self.master[k])
                full[k] = torch.cat(parts)
            else:
                full[k] = self.master[k]
        params = self.projector._ordered_params()
        for p, kind, off in zip(params, self._lay["kinds"], self._lay["offsets"]):
            p.data.copy_(full[kind][off:off + p.numel()].view(p.shape).to(p.dtype))
        self._param_versions = [(p._version, p.data_ptr()) for p in params]
        self._mark_compute_current()

    # ------------------------------------------------------------------ state (sharded, one file per rank)
    def state_dict(self) -> dict:
        self.wait()
        return {"step": self.step_count, "world": self.world, "rank": self.rank,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "master": {k: v.detach().to("cpu", copy=True) for k, v in self.master.items()},
                "exp_avg": {k: v.detach().to("cpu", copy=True) for k, v in self.exp_avg.items()},
                "exp_avg_sq": {k: v.detach().to("cpu", copy=True) for k, v in self.exp_avg_sq.items()}}

    def load_state_dict(self, state: dict) -> None:
        if state["world"] != self.world or state["rank"] != self.rank:
            raise RuntimeError("ShardedAdamW state was saved for another (world, rank)")
        self.step_count = int(state["step"])
        for g, saved in zip(self.param_groups, state["param_groups"]):
            g.update(saved)
        for name in ("master", "exp_avg", "exp_avg_sq"):
            for k in ("mat", "vec"):
                getattr(self, name)[k].copy_(state[name][k])
        # rebuild W from the masters: own slice directly, peers' slices on the next sync / step (all ranks load)
        self.sync_parameters()
        self.reload_masters()


class _CudaKernels:
    """The two kernels ShardedAdamW launches, through the C ABI (tests inject a torch restatement for gloo on CPU)."""

    def adamw_reduce_step(self, master, exp_avg, exp_avg_sq, srcs, weight_out, step, *, lr, betas, eps, weight_decay,
                          grad_scale) -> None:
        ops.adamw_reduce_step(master, exp_avg, exp_avg_sq, srcs, weight_out, step, lr=lr, betas=betas, eps=eps,
                              weight_decay=weight_decay, grad_scale=grad_scale)

    def cast_f32_bf16(self, src, dst) -> None:
        ops.cast_f32_bf16(src, dst)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
