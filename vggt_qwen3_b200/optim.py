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
                if p.numel() % 4 != 0:
                    raise RuntimeError("FusedAdamW: parameter sizes must be multiples of 4")
                st = self.state[p]
                if not st:
                    st["step"] = 0
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
