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

Train mode (`.train()` with `config.dropout > 0`, how the reference trains: src/train/train_sft.py:196) applies the
reference's four dropout sites per layer (:33 attention probabilities, :46 after the attention output projection, :37
inside the MLP, :49 after the MLP) with counter-based Philox masks: (seed, offset) are taken from torch's CUDA
generator like any ATen dropout (reproducible under `torch.manual_seed`), the backward regenerates the masks.

Buffers: the activations kept for backward, the scratch workspace and ONE flat f32 gradient buffer (`grad_flat`,
`p.grad` of every parameter is a view into it) persist across steps; gradient accumulation (`p.grad` still set when
the next backward runs, as inside `accelerator.accumulate`) is done by the kernels directly into that buffer.
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
        self._grad_view_list: Optional[List[torch.Tensor]] = None
        self._flat_weights = None       # (bf16 matrices, f32 vectors) flat compute-copy storage, see use_flat_storage
        self._flat_grads = None         # (matrices, vectors, views) flat gradient storage owned by an optimizer
        self._flat_grads_fresh = True   # next backward overwrites the owned gradient storage (else accumulates)
        # called (no arguments) at the top of every forward: lets an asynchronous optimizer make the current stream
        # wait for the weight update / all-gather it still has in flight
        self.before_forward: Optional[Callable[[], None]] = None
        self._scratch_bufs: dict = {}   # (kind, nbytes, device) -> persistent uint8 scratch
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
            if self._flat_weights is not None:      # caller-owned flat storage (latents: bf16 copy next to it)
                if self._flat_weights[0].device != device:
                    raise RuntimeError("flat compute-copy storage lives on another device")
                bufs = self._views_into(*self._flat_weights)
                bufs[0] = torch.empty(params[0].shape, device=device, dtype=torch.bfloat16)
            else:
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

    def flat_layout(self) -> dict:
        """Two-region flat layout shared by the gradient buffer, the compute copies and the sharded optimizer:
        region "mat" = the Linear / in_proj weight matrices (what the wgrad GEMMs produce; bf16 on the wire), region
        "vec" = biases, LayerNorm affine and the latents (f32 everywhere).  Every tensor starts at a multiple of 64
        elements.  Returns {"kinds": [...], "offsets": [...], "mat": total, "vec": total} in `_ordered_params` order."""
        kinds, offs, tot = [], [], {"mat": 0, "vec": 0}
        for p in self._ordered_params():
            kind = "mat" if (p.dim() == 2 and p is not self.latents) else "vec"
            kinds.append(kind)
            offs.append(tot[kind])
            tot[kind] += (p.numel() + 63) // 64 * 64
        return {"kinds": kinds, "offsets": offs, "mat": tot["mat"], "vec": tot["vec"]}

    def _views_into(self, mat: torch.Tensor, vec: torch.Tensor) -> List[torch.Tensor]:
        lay = self.flat_layout()
        out = []
        for p, kind, off in zip(self._ordered_params(), lay["kinds"], lay["offsets"]):
            src = mat if kind == "mat" else vec
            out.append(src[off:off + p.numel()].view(p.shape))
        return out

    def _grad_views(self, device: torch.device, fresh: bool = False) -> List[torch.Tensor]:
        """Views, in `_ordered_params` order, into the flat f32 gradient buffer ([mat region | vec region]).  The
        buffer persists across steps (`grad_flat`); `fresh=True` returns views into a new temporary buffer instead
        (used when some `p.grad` is set but does not alias the persistent buffer, so that autograd accumulates the
        ordinary way)."""
        lay = self.flat_layout()
        total = lay["mat"] + lay["vec"]
        if not fresh and self._grad_flat is not None and self._grad_flat.device == device \
                and self._grad_flat.numel() == total:
            return self._grad_view_list
        flat = torch.empty(total, device=device, dtype=torch.float32)
        views = self._views_into(flat[:lay["mat"]], flat[lay["mat"]:])
        if not fresh:
            self._grad_flat, self._grad_view_list = flat, views
        return views

    def use_flat_storage(self, weights_mat: torch.Tensor, weights_vec: torch.Tensor,
                         grads_mat: Optional[torch.Tensor] = None, grads_vec: Optional[torch.Tensor] = None) -> None:
        """Place the compute copies (bf16 matrices / f32 vectors) and, optionally, the gradients in caller-owned flat
        buffers laid out as `flat_layout()` -- e.g. NVLink-symmetric memory owned by optim.ShardedAdamW, which then
        updates the compute copies itself and all-gathers them between ranks.  With gradient buffers (grads_mat bf16
        or f32, grads_vec f32) the backward writes there and hands NO gradient tensors to autograd (`p.grad` stays
        None): the owner of the buffers consumes them."""
        lay = self.flat_layout()
        assert weights_mat.dtype == torch.bfloat16 and weights_vec.dtype == torch.float32
        assert weights_mat.numel() >= lay["mat"] and weights_vec.numel() >= lay["vec"]
        self._flat_weights = (weights_mat, weights_vec)
        self._compute = None            # rebuilt on the next forward as views into the flat buffers
        if grads_mat is not None:
            assert grads_mat.dtype in (torch.bfloat16, torch.float32) and grads_vec.dtype == torch.float32
            assert grads_mat.numel() >= lay["mat"] and grads_vec.numel() >= lay["vec"]
            self._flat_grads = (grads_mat, grads_vec, self._views_into(grads_mat, grads_vec))
        else:
            self._flat_grads = None
        self._flat_grads_fresh = True

    def mark_grads_consumed(self) -> None:
        """Owner of the flat gradient storage: the next backward overwrites it instead of accumulating."""
        self._flat_grads_fresh = True

    @property
    def grad_flat(self) -> Optional[torch.Tensor]:
        """The flat f32 gradient buffer of the last backward (None before the first)."""
        return self._grad_flat

    def _scratch(self, kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
        """Persistent 1024-byte-aligned scratch (`saved` activations / workspace), reused while the size matches."""
        buf = self._scratch_bufs.get(kind)
        if buf is None or buf.numel() < nbytes + 1024 or buf.device != device:
            buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._scratch_bufs[kind] = buf
        return buf

    def _lease_saved(self, nbytes: int, device: torch.device) -> "_Lease":
        """A `saved`-activations buffer from the module's pool (stable addresses across steps: the TMA descriptors
        cached by the library keep hitting); it goes back to the pool when the autograd graph that holds it dies."""
        pool = self._scratch_bufs.setdefault("saved_pool", [])
        for i, buf in enumerate(pool):
            if buf.numel() >= nbytes + 1024 and buf.device == device:
                return _Lease(pool, pool.pop(i))
        del pool[:]     # shape changed: drop stale buffers
        return _Lease(pool, torch.empty(nbytes + 1024, dtype=torch.uint8, device=device))

    def _draw_dropout(self, device: torch.device):
        """(p, seed, offset) of this forward: eval mode or dropout 0 -> (0, 0, 0); train mode advances the CUDA
        generator's Philox offset exactly like an ATen dropout kernel would."""
        p = float(self.config.dropout)
        if not self.training or p <= 0.0:
            return 0.0, 0, 0
        if not 0.0 <= p < 1.0:
            raise ValueError(f"dropout probability must be in [0, 1), got {p}")
        gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        seed, offset = int(gen.initial_seed()), int(gen.get_offset())
        gen.set_offset(offset + 4)
        return p, seed & 0xFFFFFFFFFFFFFFFF, offset

    def _cfg(self, B: int, T: int, out_dtype: torch.dtype, dropout=(0.0, 0, 0)) -> PerceiverCfg:
        c = PerceiverCfg()
        c.dropout_p, c.dropout_seed, c.dropout_offset = dropout
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
        if self.before_forward is not None:
            self.before_forward()
        return _PerceiverFunction.apply(self, tokens, *self._ordered_params())


class _Lease:
    """Returns `buf` to `pool` when the last reference (the autograd ctx) goes away."""

    def __init__(self, pool: list, buf: torch.Tensor) -> None:
        self.pool, self.buf = pool, buf

    def __del__(self) -> None:
        if len(self.pool) < 4:
            self.pool.append(self.buf)


class _PerceiverFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: PerceiverProjector, tokens: torch.Tensor, *params: torch.Tensor) -> torch.Tensor:
        lib = _lib.lib()
        dev = tokens.device
        B, T, _ = tokens.shape
        out_dtype = torch.float32 if module.in_proj.weight.dtype == torch.float32 else torch.bfloat16
        cfg = module._cfg(B, T, out_dtype, module._draw_dropout(dev))
        cc = module._refresh_compute_copies(dev)
        tok = tokens.detach().reshape(B * T, -1)
        tok = tok.contiguous() if tok.dtype == torch.bfloat16 else tok.to(torch.bfloat16)
        saved_bytes = lib.vq_perceiver_saved_bytes(C.byref(cfg))
        work_bytes = lib.vq_perceiver_workspace_bytes(C.byref(cfg))
        if saved_bytes == 0 or work_bytes == 0:
            _lib.check(-1, "vq_perceiver_saved_bytes")
        lease = module._lease_saved(saved_bytes, dev)   # returns to the module's pool when the graph is freed
        saved = lease.buf
        work = module._scratch("work", work_bytes, dev)
        out = torch.empty(B, module.config.num_latents, module.out_dim, dtype=out_dtype, device=dev)
        a1024 = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
        rc = lib.vq_perceiver_forward(C.byref(cfg), C.byref(cc["structs"][0]), tok.data_ptr(), out.data_ptr(),
                                      a1024(saved), a1024(work), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vq_perceiver_forward")
        ctx.module, ctx.cfg, ctx.saved, ctx.tok, ctx.work, ctx.lease = module, cfg, saved, tok, work, lease
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
        params = module._ordered_params()
        # Gradient hand-off.  `grad_flat` persists; three cases:
        #  (a) every p.grad is None (zero_grad(set_to_none=True), first step): the kernels OVERWRITE the flat buffer and
        #      the views are returned to autograd, which adopts them as p.grad without a copy;
        #  (b) every p.grad IS the matching view (gradient accumulation, zero_grad(set_to_none=False)): the kernels
        #      ACCUMULATE into the flat buffer in place and autograd gets no gradient tensors (nothing to add);
        #  (c) anything else (a foreign p.grad): a temporary buffer, ordinary autograd accumulation.
        #  (d) flat gradient storage owned by an optimizer (use_flat_storage): written / accumulated there, matrices
        #      possibly in bf16; autograd gets no gradient tensors and `p.grad` stays None.
        owned = module._flat_grads is not None
        if owned:
            views = module._flat_grads[2]
            in_place = True
            accumulate = 0 if module._flat_grads_fresh else 1
            module._flat_grads_fresh = False
        else:
            views = module._grad_views(dev)
            has_grad = [p.grad is not None for p in params]
            aliased = all(h and p.grad.data_ptr() == v.data_ptr() and p.grad.dtype == torch.float32
                          for h, p, v in zip(has_grad, params, views))
            in_place = aliased and any(has_grad)
            if any(has_grad) and not in_place:
                views = module._grad_views(dev, fresh=True)
            accumulate = 1 if in_place else 0
        n = len(_LAYER_PARAM_ORDER)
        gr = PerceiverGrads()
        gr.latents, gr.in_proj_w, gr.in_proj_b = views[0].data_ptr(), views[1].data_ptr(), views[2].data_ptr()
        arr = (PerceiverLayerGrads * cfg.layers)()
        for l in range(cfg.layers):
            for j, (_, field, _) in enumerate(_LAYER_PARAM_ORDER):
                setattr(arr[l], field, views[3 + l * n + j].data_ptr())
        gr.layers = arr
        gr.out_proj_w, gr.out_proj_b = views[-2].data_ptr(), views[-1].data_ptr()
        gr.accumulate = accumulate
        gr.weight_grad_bf16 = 1 if (owned and module._flat_grads[0].dtype == torch.bfloat16) else 0
        need_tok = ctx.needs_input_grad[1]
        grad_tok = torch.empty(cfg.B * cfg.T, cfg.in_dim, dtype=torch.float32, device=dev) if need_tok else None
        hook = module.grad_ready_hook
        hook_error: List[BaseException] = []

        def on_layer_done(layer, _user):   # a ctypes callback must not raise: stash the exception, re-raise below
            if hook is None or hook_error:
                return
            try:
                hook(int(layer))
            except BaseException as exc:  # noqa: BLE001
                hook_error.append(exc)

        cb = LayerDoneFn(on_layer_done)
        a1024 = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
        rc = lib.vq_perceiver_backward(
            C.byref(cfg), C.byref(cc["structs"][0]), ctx.tok.data_ptr(), g.data_ptr(), a1024(ctx.saved), C.byref(gr),
            grad_tok.data_ptr() if need_tok else None, a1024(ctx.work), cb, None,
            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vq_perceiver_backward")
        if hook_error:
            raise RuntimeError("grad_ready_hook failed during the Perceiver backward") from hook_error[0]
        if in_place:
            grads = [None] * len(params)
        else:
            # fresh tensor objects over the same storage: autograd adopts a gradient without copying only when it
            # holds the sole reference to it (the module keeps its own view objects)
            grads = [v.view(p.shape) if p.dtype == torch.float32 else v.to(p.dtype) for v, p in zip(views, params)]
        gt = None
        if need_tok:
            gt = grad_tok.view(ctx.tokens_shape).to(ctx.tokens_dtype)
        ctx.saved = ctx.work = ctx.tok = ctx.lease = None
        return (None, gt, *grads)
