"""Perceiver projector with the reference's module interface, executing on libvq_b200.so.

Mirrors src/models/projector_perceiver.py of the reference: `PerceiverConfig` (:20-27, same field names as
configs/perceiver_small.yaml), `PerceiverLayer` (:30-50), `PerceiverProjector` (:53-82) — same constructor
signatures, same parameter names (so `projector.*` checkpoint keys are unchanged: latents, in_proj.*,
layers.N.self_attn.{in_proj_weight,in_proj_bias,out_proj.*}, layers.N.mlp.{0,3}.*, layers.N.norm{1,2}.*,
out_proj.*), same initialisation (`_init_weights`, :13-17).

The torch.nn submodules are parameter containers; `forward` runs `vq_perceiver_forward` (tcgen05 GEMMs + row
kernels) and autograd is provided by `_PerceiverFunction`, whose backward is `vq_perceiver_backward`.
f32 master parameters get bf16 compute copies (refreshed when a parameter's version counter changes);
gradients come back in f32.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import (VQ_BF16, VQ_F32, LayerDoneFn, PerceiverCfg, PerceiverGrads, PerceiverLayerGrads,
                   PerceiverLayerWeights, PerceiverWeights)


@dataclass
class PerceiverConfig:
    latent_dim: int = 4096
    num_latents: int = 128
    num_heads: int = 8
    num_layers: int = 6
    ffn_dim: int = 16384
    dropout: float = 0.1


def _xavier_linear_init(module: nn.Module) -> None:
    # reference `_init_weights`: xavier-uniform weights and zero biases for every nn.Linear (incl. MHA out_proj)
    if isinstance(module, (nn.Linear, nn.Conv1d)):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.zeros_(module.bias)


class PerceiverLayer(nn.Module):
    """Parameter container for one cross-attention + MLP layer (post-LayerNorm), reference :30-50."""

    def __init__(self, dim: int, heads: int, ffn_dim: int, dropout: float) -> None:
        super().__init__()
        self.self_attn = nn.MultiheadAttention(dim, heads, dropout=dropout, batch_first=True)
        self.mlp = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(), nn.Dropout(dropout), nn.Linear(ffn_dim, dim))
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, latents: torch.Tensor, context: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise RuntimeError("PerceiverLayer is executed by PerceiverProjector.forward through the C ABI as a whole")


_LAYER_PARAM_ORDER = (
    ("self_attn.in_proj_weight", "in_proj_w", True), ("self_attn.in_proj_bias", "in_proj_b", False),
    ("self_attn.out_proj.weight", "out_proj_w", True), ("self_attn.out_proj.bias", "out_proj_b", False),
    ("norm1.weight", "norm1_w", False), ("norm1.bias", "norm1_b", False),
    ("mlp.0.weight", "fc1_w", True), ("mlp.0.bias", "fc1_b", False),
    ("mlp.3.weight", "fc2_w", True), ("mlp.3.bias", "fc2_b", False),
    ("norm2.weight", "norm2_w", False), ("norm2.bias", "norm2_b", False),
)


class PerceiverProjector(nn.Module):
    """Resample VGGT aggregated tokens [B,T,in_dim] to [B,num_latents,out_dim] (reference :53-82)."""

    def __init__(self, config: PerceiverConfig, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.config = config
        self.in_dim, self.out_dim = in_dim, out_dim
        self.latents = nn.Parameter(torch.randn(config.num_latents, config.latent_dim) * 0.02)
        self.in_proj = nn.Linear(in_dim, config.latent_dim)
        self.layers = nn.ModuleList(
            [PerceiverLayer(config.latent_dim, config.num_heads, config.ffn_dim, config.dropout)
             for _ in range(config.num_layers)])
        self.out_proj = nn.Linear(config.latent_dim, out_dim)
        self.apply(_xavier_linear_init)
        self._compute: Optional[dict] = None
        self._grad_flat: Optional[torch.Tensor] = None
        # called as hook(layer_index) right after the kernels producing that bucket's gradients are enqueued
        # (layer_index = num_layers for out_proj, L-1..0 for the layers, -1 for in_proj + latents)
        self.grad_ready_hook: Optional[Callable[[int], None]] = None

    # ------------------------------------------------------------------ parameters
    def _ordered_params(self) -> List[nn.Parameter]:
        """Fixed order shared by the autograd function and the flat gradient buffer."""
        ps = [self.latents, self.in_proj.weight, self.in_proj.bias]
        for layer in self.layers:
            named = dict(layer.named_parameters())
            ps += [named[n] for n, _, _ in _LAYER_PARAM_ORDER]
        ps += [self.out_proj.weight, self.out_proj.bias]
        return ps

    def grad_buckets(self) -> List[List[nn.Parameter]]:
        """Parameters grouped by the backward's completion order: out_proj, layers L-1..0, in_proj+latents."""
        buckets = [[self.out_proj.weight, self.out_proj.bias]]
        for layer in reversed(self.layers):
            named = dict(layer.named_parameters())
            buckets.append([named[n] for n, _, _ in _LAYER_PARAM_ORDER])
        buckets.append([self.latents, self.in_proj.weight, self.in_proj.bias])
        return buckets

    def _refresh_compute_copies(self, device: torch.device) -> dict:
        """bf16 copies of matrices + latents, f32 copies of vectors; re-cast only tensors whose version changed."""
        params = self._ordered_params()
        cc = self._compute
        if cc is None or cc["device"] != device or cc["ids"] != [id(p) for p in params]:
            is_mat = [p.dim() == 2 for p in params]
            bufs = [torch.empty(p.shape, device=device, dtype=torch.bfloat16 if m else torch.float32)
                    for p, m in zip(params, is_mat)]
            cc = {"device": device, "ids": [id(p) for p in params], "bufs": bufs, "versions": [-1] * len(params),
                  "ptrs": [None] * len(params), "structs": None}
            self._compute = cc
        lib = _lib.lib()
        stream = torch.cuda.current_stream().cuda_stream
        for i, p in enumerate(params):
            if p.device != device:
                raise RuntimeError("PerceiverProjector parameters must live on the CUDA device of the input")
            ver = (p._version, p.data_ptr())
            if cc["versions"][i] == ver:
                continue
            buf = cc["bufs"][i]
            if p.dtype == buf.dtype:
                buf.copy_(p.detach())
            elif p.dtype == torch.float32 and buf.dtype == torch.bfloat16 and p.is_contiguous() and p.data_ptr() % 16 == 0:
                _lib.check(lib.vq_cast_f32_bf16(p.data_ptr(), buf.data_ptr(), p.numel(), stream), "vq_cast_f32_bf16")
            else:
                buf.copy_(p.detach())
            cc["versions"][i] = ver
        if cc["structs"] is None:
            bufs = cc["bufs"]
            w = PerceiverWeights()
            w.latents, w.in_proj_w, w.in_proj_b = bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr()
            n = len(_LAYER_PARAM_ORDER)
            arr = (PerceiverLayerWeights * len(self.layers))()
            for l in range(len(self.layers)):
                for j, (_, field, _) in enumerate(_LAYER_PARAM_ORDER):
                    setattr(arr[l], field, bufs[3 + l * n + j].data_ptr())
            w.layers = arr
            w.out_proj_w, w.out_proj_b = bufs[-2].data_ptr(), bufs[-1].data_ptr()
            cc["structs"] = (w, arr)
        return cc

    def _grad_views(self, device: torch.device) -> List[torch.Tensor]:
        """Views, in `_ordered_params` order, into one flat f32 gradient buffer (fresh per backward so that
        autograd can take ownership of the views without aliasing a previous step's .grad)."""
        params = self._ordered_params()
        total = sum((p.numel() + 63) // 64 * 64 for p in params)
        flat = torch.empty(total, device=device, dtype=torch.float32)
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view(p.shape))
            off += (p.numel() + 63) // 64 * 64
        self._grad_flat = flat
        return views

    def _cfg(self, B: int, T: int, out_dtype: torch.dtype) -> PerceiverCfg:
        c = PerceiverCfg()
        cfg = self.config
        c.B, c.T, c.in_dim = B, T, self.in_dim
        c.E, c.heads, c.Lq, c.layers, c.ffn = cfg.latent_dim, cfg.num_heads, cfg.num_latents, cfg.num_layers, cfg.ffn_dim
        c.out_dim = self.out_dim
        c.ln_eps = float(self.layers[0].norm1.eps)
        c.out_dtype = VQ_F32 if out_dtype == torch.float32 else VQ_BF16
        return c

    # ------------------------------------------------------------------ execution
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens [B, T, in_dim] -> [B, num_latents, out_dim] in the parameters' dtype."""
        if tokens.dim() != 3 or tokens.shape[-1] != self.in_dim:
            raise ValueError(f"expected tokens [B, T, {self.in_dim}], got {tuple(tokens.shape)}")
        if not tokens.is_cuda:
            raise RuntimeError("PerceiverProjector runs on sm_100a only: tokens must be a CUDA tensor (no CPU fallback)")
        if self.training and self.config.dropout > 0.0:
            raise NotImplementedError(
                "dropout > 0 in training mode is not implemented in the sm_100a kernels yet; "
                "call .eval() or build the projector with PerceiverConfig(dropout=0.0)")
        return _PerceiverFunction.apply(self, tokens, *self._ordered_params())


class _PerceiverFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: PerceiverProjector, tokens: torch.Tensor, *params: torch.Tensor) -> torch.Tensor:
        lib = _lib.lib()
        dev = tokens.device
        B, T, _ = tokens.shape
        out_dtype = torch.float32 if module.in_proj.weight.dtype == torch.float32 else torch.bfloat16
        cfg = module._cfg(B, T, out_dtype)
        cc = module._refresh_compute_copies(dev)
        tok = tokens.detach().reshape(B * T, -1)
        tok = tok.contiguous() if tok.dtype == torch.bfloat16 else tok.to(torch.bfloat16)
        saved_bytes = lib.vq_perceiver_saved_bytes(C.byref(cfg))
        work_bytes = lib.vq_perceiver_workspace_bytes(C.byref(cfg))
        if saved_bytes == 0 or work_bytes == 0:
            _lib.check(-1, "vq_perceiver_saved_bytes")
        saved = torch.empty(saved_bytes + 1024, dtype=torch.uint8, device=dev)
        work = torch.empty(work_bytes + 1024, dtype=torch.uint8, device=dev)
        out = torch.empty(B, module.config.num_latents, module.out_dim, dtype=out_dtype, device=dev)
        a1024 = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
        rc = lib.vq_perceiver_forward(C.byref(cfg), C.byref(cc["structs"][0]), tok.data_ptr(), out.data_ptr(),
                                      a1024(saved), a1024(work), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vq_perceiver_forward")
        ctx.module, ctx.cfg, ctx.saved, ctx.tok, ctx.work = module, cfg, saved, tok, work
        ctx.tokens_dtype = tokens.dtype
        ctx.tokens_shape = tokens.shape
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        lib = _lib.lib()
        module: PerceiverProjector = ctx.module
        cfg: PerceiverCfg = ctx.cfg
        dev = grad_out.device
        cc = module._refresh_compute_copies(dev)
        g = grad_out.detach().reshape(cfg.B * cfg.Lq, cfg.out_dim)
        g = g.contiguous() if g.dtype == torch.bfloat16 else g.to(torch.bfloat16)
        views = module._grad_views(dev)
        n = len(_LAYER_PARAM_ORDER)
        gr = PerceiverGrads()
        gr.latents, gr.in_proj_w, gr.in_proj_b = views[0].data_ptr(), views[1].data_ptr(), views[2].data_ptr()
        arr = (PerceiverLayerGrads * cfg.layers)()
        for l in range(cfg.layers):
            for j, (_, field, _) in enumerate(_LAYER_PARAM_ORDER):
                setattr(arr[l], field, views[3 + l * n + j].data_ptr())
        gr.layers = arr
        gr.out_proj_w, gr.out_proj_b = views[-2].data_ptr(), views[-1].data_ptr()
        gr.accumulate = 0
        need_tok = ctx.needs_input_grad[1]
        grad_tok = torch.empty(cfg.B * cfg.T, cfg.in_dim, dtype=torch.float32, device=dev) if need_tok else None
        hook = module.grad_ready_hook
        cb = LayerDoneFn((lambda layer, _user: hook(int(layer))) if hook is not None else (lambda layer, _user: None))
        a1024 = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
        rc = lib.vq_perceiver_backward(
            C.byref(cfg), C.byref(cc["structs"][0]), ctx.tok.data_ptr(), g.data_ptr(), a1024(ctx.saved), C.byref(gr),
            grad_tok.data_ptr() if need_tok else None, a1024(ctx.work), cb, None,
            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vq_perceiver_backward")
        params = module._ordered_params()
        grads = [v if p.dtype == torch.float32 else v.to(p.dtype) for v, p in zip(views, params)]
        gt = None
        if need_tok:
            gt = grad_tok.view(ctx.tokens_shape).to(ctx.tokens_dtype)
        ctx.saved = ctx.work = ctx.tok = None
        return (None, gt, *grads)
