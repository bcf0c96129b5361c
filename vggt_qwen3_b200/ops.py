"""Thin torch-tensor wrappers over the C ABI (include/vq_b200.h).

PyTorch is used only for device memory and streams: every function takes CUDA tensors, passes raw
pointers to libvq_b200.so on the current stream and returns the output tensor.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EPI_GELU, EPI_GELU_BWD, EPI_LINEAR, EPI_PATCH, EPI_QKV_ROPE, VQ_BF16, VQ_F32, GemmDesc


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


def _dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return VQ_BF16
    if dtype == torch.float32:
        return VQ_F32
    raise RuntimeError(f"unsupported output dtype {dtype}")


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    a_major: int = 0,
    b_major: int = 0,
    out: Optional[torch.Tensor] = None,
    out_dtype: torch.dtype = torch.bfloat16,
    epilogue: int = EPI_LINEAR,
    alpha: float = 1.0,
    bias: Optional[torch.Tensor] = None,
    gamma: Optional[torch.Tensor] = None,
    add: Optional[torch.Tensor] = None,
    aux: Optional[torch.Tensor] = None,
    extra: Optional[dict] = None,
) -> torch.Tensor:
    """C = epilogue(A @ B^T) for 2-D bf16 operands.

    a_major=0: a is [M,K]; a_major=1: a is [K,M] (M contiguous).  Likewise b: [N,K] or [K,N].
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_major else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_major else (b.shape[0], b.shape[1])
    if K != Kb:
        raise RuntimeError(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_major, d.b_major = a_major, b_major
    d.A, d.lda = a.data_ptr(), a.stride(0)
    d.B, d.ldb = b.data_ptr(), b.stride(0)
    d.C, d.ldc = out.data_ptr(), out.stride(0)
    d.c_dtype = _dtype_code(out.dtype)
    d.Z0 = d.Z1 = 1
    d.epilogue = epilogue
    d.alpha = alpha
    if bias is not None:
        _req(bias, torch.float32, "bias")
        d.bias = bias.data_ptr()
    if gamma is not None:
        _req(gamma, torch.float32, "gamma")
        d.gamma = gamma.data_ptr()
    if add is not None:
        _req(add, out.dtype, "add")
        d.add, d.ldadd = add.data_ptr(), add.stride(0)
    if aux is not None:
        _req(aux, torch.bfloat16, "aux")
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    if extra:
        for k, v in extra.items():
            setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    _lib.check(_lib.lib().vq_gemm(C.byref(d), _stream()), "vq_gemm")
    return out


def gemm_tune(tail_split_max: int = -1) -> None:
    """Maximum K slices per tail tile of the persistent GEMM schedule (0 / 1 = no split); -1 keeps the setting."""
    _lib.check(_lib.lib().vq_gemm_tune(tail_split_max), "vq_gemm_tune")


def gemm_desc(**kw) -> None:
    """Launch a GEMM from explicit descriptor fields (batched / strided use)."""
    d = GemmDesc()
    for k, v in kw.items():
        setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    _lib.check(_lib.lib().vq_gemm(C.byref(d), _stream()), "vq_gemm")


def attention_fwd(qkv: torch.Tensor, n_seq: int, seq_len: int, heads: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(QK^T)V over a packed [n_seq*seq_len, 3*heads*64] bf16 buffer (q pre-scaled by log2e/8)."""
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.dim() == 2 and qkv.is_contiguous() and qkv.shape == (n_seq * seq_len, 3 * heads * 64)
    if out is None:
        out = torch.empty((n_seq * seq_len, heads * 64), device=qkv.device, dtype=torch.bfloat16)
    _lib.check(
        _lib.lib().vq_attention_fwd(qkv.data_ptr(), out.data_ptr(), n_seq, seq_len, heads, _stream()),
        "vq_attention_fwd",
    )
    return out


def attention_fwd_bounded(qkv: torch.Tensor, n_seq: int, seq_len: int, heads: int, score_bound: float,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`attention_fwd` for buffers whose scores are PROVEN to satisfy |q . k| <= score_bound <= 64 (see the header)."""
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.dim() == 2 and qkv.is_contiguous() and qkv.shape == (n_seq * seq_len, 3 * heads * 64)
    if out is None:
        out = torch.empty((n_seq * seq_len, heads * 64), device=qkv.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().vq_attention_fwd_bounded(qkv.data_ptr(), out.data_ptr(), n_seq, seq_len, heads,
                                                   float(score_bound), _stream()), "vq_attention_fwd_bounded")
    return out


def attention_tune_bounded(enable: int = -1, emu_pairs: int = -1) -> None:
    _lib.check(_lib.lib().vq_attention_tune_bounded(enable, emu_pairs), "vq_attention_tune_bounded")


def attention_tune(q_tiles_per_cta: int = -1, emu_pairs: int = -1) -> None:
    """Force an attention kernel variant (tests / tuning); -1 keeps the current setting, q_tiles_per_cta=0 = auto."""
    _lib.check(_lib.lib().vq_attention_tune(q_tiles_per_cta, emu_pairs), "vq_attention_tune")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out_dtype: torch.dtype = torch.bfloat16, return_stats: bool = False):
    """Row LayerNorm of a 2-D f32/bf16 tensor (D % 128 == 0)."""
    assert x.dim() == 2 and x.is_contiguous() and x.is_cuda
    rows, D = x.shape
    out = torch.empty(rows, D, device=x.device, dtype=out_dtype)
    mean = rstd = None
    if return_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    _lib.check(
        _lib.lib().vq_layernorm(x.data_ptr(), _dtype_code(x.dtype), out.data_ptr(), _dtype_code(out_dtype),
                                gamma.data_ptr(), beta.data_ptr(), rows, D, eps, _ptr(mean), _ptr(rstd), _stream()),
        "vq_layernorm")
    return (out, mean, rstd) if return_stats else out


def dropout_desc(p: float, seed: int, offset: int, site: int) -> "_lib.DropoutDesc":
    d = _lib.DropoutDesc()
    d.p, d.site, d.seed, d.offset = float(p), int(site), int(seed), int(offset)
    return d


def dropout_mask(rows: int, cols: int, ld: int, p: float, seed: int, offset: int, site: int,
                 device: Optional[torch.device] = None) -> torch.Tensor:
    """uint8 [rows, cols] keep-mask of a dropout site (test hook: the mask the kernels regenerate on the fly)."""
    keep = torch.empty(rows, cols, dtype=torch.uint8, device=device or torch.device("cuda"))
    d = dropout_desc(p, seed, offset, site)
    _lib.check(_lib.lib().vq_dropout_mask(keep.data_ptr(), rows, cols, ld, C.byref(d), _stream()), "vq_dropout_mask")
    return keep


def perceiver_attention_supported(num_latents: int, T: int, latent_dim: int, heads: int) -> bool:
    """True when the Perceiver's attention core runs as the one-tile kernels (csrc/perceiver_attn.cu)."""
    return bool(_lib.lib().vq_perceiver_attention_supported(num_latents, T, latent_dim, heads))


def perceiver_attention_tune(fused: int = -1) -> None:
    """0: vq_perceiver_forward / _backward take the multi-launch formulation (A/B, tests); 1: one-tile kernels."""
    _lib.check(_lib.lib().vq_perceiver_attention_tune(fused), "vq_perceiver_attention_tune")


def perceiver_attention_fwd(q: torch.Tensor, kv: torch.Tensor, B: int, heads: int,
                            dropout: Optional["_lib.DropoutDesc"] = None):
    """nn.MultiheadAttention's core for q [B*Lq, E], kv [B*T, 2E] = k | v (bf16): returns (att [B*Lq, E], probs
    [B*heads*Lq, Tp], probs_dropped or None); one launch, scores stay in TMEM."""
    _req(q, torch.bfloat16, "q")
    _req(kv, torch.bfloat16, "kv")
    assert q.dim() == 2 and kv.dim() == 2 and q.is_contiguous() and kv.is_contiguous()
    E = q.shape[1]
    assert kv.shape[1] == 2 * E and q.shape[0] % B == 0 and kv.shape[0] % B == 0
    Lq, T = q.shape[0] // B, kv.shape[0] // B
    Tp = (T + 63) // 64 * 64
    probs = torch.empty(B * heads * Lq, Tp, device=q.device, dtype=torch.bfloat16)
    dropped = torch.empty_like(probs) if dropout is not None and dropout.p > 0 else None
    att = torch.empty_like(q)
    _lib.check(_lib.lib().vq_perceiver_attention_fwd(q.data_ptr(), kv.data_ptr(), probs.data_ptr(), _ptr(dropped),
                                                     att.data_ptr(), B, heads, Lq, T, E,
                                                     C.byref(dropout) if dropout is not None else None, _stream()),
               "vq_perceiver_attention_fwd")
    return att, probs, dropped


def perceiver_attention_bwd(q: torch.Tensor, kv: torch.Tensor, probs: torch.Tensor, datt: torch.Tensor, B: int,
                            heads: int, dropout: Optional["_lib.DropoutDesc"] = None):
    """Backward of `perceiver_attention_fwd`: (dq [B*Lq, E], dkv [B*T, 2E] = dk | dv), bf16, one launch."""
    for t, n in ((q, "q"), (kv, "kv"), (probs, "probs"), (datt, "datt")):
        _req(t, torch.bfloat16, n)
        assert t.is_contiguous()
    E = q.shape[1]
    Lq, T = q.shape[0] // B, kv.shape[0] // B
    assert datt.shape == q.shape and probs.shape == (B * heads * Lq, (T + 63) // 64 * 64)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    _lib.check(_lib.lib().vq_perceiver_attention_bwd(q.data_ptr(), kv.data_ptr(), probs.data_ptr(), datt.data_ptr(),
                                                     dq.data_ptr(), dkv.data_ptr(), B, heads, Lq, T, E,
                                                     C.byref(dropout) if dropout is not None else None, _stream()),
               "vq_perceiver_attention_bwd")
    return dq, dkv


def row_stats_cast(x: torch.Tensor, stats_zero: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(xb = bf16(x), stats f32 [rows, 2] = per-row (sum, sum of squares)) of an f32 [rows, D] tensor: the entry of a
    LayerNorm-folded GEMM chain (vq_gemm_desc.ln_stats); `stats_zero` (f32 [rows, 2]) is cleared in the same launch."""
    _req(x, torch.float32, "x")
    assert x.dim() == 2 and x.is_contiguous()
    rows, D = x.shape
    xb = torch.empty(rows, D, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty(rows, 2, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().vq_row_stats_cast(x.data_ptr(), xb.data_ptr(), stats.data_ptr(), _ptr(stats_zero), rows, D,
                                            _stream()), "vq_row_stats_cast")
    return xb, stats


def layernorm_bwd(dy: torch.Tensor, z: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                  dropout: Optional["_lib.DropoutDesc"] = None):
    """Returns (dz bf16, dgamma f32, dbeta f32) for bf16 dy and saved LN input z; with `dropout` additionally the
    masked + rescaled dz of the branch that went through dropout before the residual add."""
    _req(dy, torch.bfloat16, "dy")
    _req(z, torch.bfloat16, "z")
    rows, D = z.shape
    dz = torch.empty_like(z)
    dgamma = torch.zeros(D, device=z.device, dtype=torch.float32)
    dbeta = torch.zeros(D, device=z.device, dtype=torch.float32)
    dzd = torch.empty_like(z) if dropout is not None else None
    _lib.check(
        _lib.lib().vq_layernorm_bwd(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                    dz.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), rows, D,
                                    C.byref(dropout) if dropout is not None else None, _ptr(dzd), _stream()),
        "vq_layernorm_bwd")
    return (dz, dgamma, dbeta) if dropout is None else (dz, dgamma, dbeta, dzd)


def splice_overwrite(inputs_embeds: torch.Tensor, input_ids: torch.Tensor, features: torch.Tensor, image_id: int,
                     overflow: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place training splice: inputs_embeds[b, pos:pos+n] = features[b] for every <image> position pos.

    Mirrors src/models/vggt_qwen3_vlm.py:190-195 without the host `.nonzero()` sync.  `overflow` (optional int32
    device scalar) counts spans that do not fit in the sequence (the reference raises in that case)."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("splice_overwrite: CUDA tensors only")
    B, L, H = inputs_embeds.shape
    assert inputs_embeds.is_contiguous() and input_ids.shape == (B, L) and features.shape[0] == B and features.shape[2] == H
    ids = input_ids.to(torch.int64).contiguous()
    feats = features.to(inputs_embeds.dtype).contiguous()
    _lib.check(
        _lib.lib().vq_splice_overwrite(inputs_embeds.data_ptr(), _dtype_code(inputs_embeds.dtype), ids.data_ptr(),
                                       feats.data_ptr(), B, L, H, feats.shape[1], int(image_id), _ptr(overflow),
                                       _stream()),
        "vq_splice_overwrite")
    return inputs_embeds


def splice_features(inputs_embeds: torch.Tensor, input_ids: torch.Tensor, geom: Optional[torch.Tensor], n_geom: int,
                    vis: torch.Tensor, image_id: int, overflow: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place training splice of `cat([geom.expand(n_geom), vis], 1)` over every <image> position without building
    the span (src/models/vggt_qwen3_vlm.py:187-195): geom f32 [B, hidden] (or None), vis [B, n_vis, hidden]."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("splice_features: CUDA tensors only")
    B, L, H = inputs_embeds.shape
    assert inputs_embeds.is_contiguous() and input_ids.shape == (B, L) and vis.shape[0] == B and vis.shape[2] == H
    ids = input_ids.to(torch.int64).contiguous()
    v = vis.to(inputs_embeds.dtype).contiguous()
    g = None
    if geom is not None and n_geom > 0:
        assert geom.shape == (B, H)
        g = geom.to(torch.float32).contiguous()
    _lib.check(
        _lib.lib().vq_splice_features(inputs_embeds.data_ptr(), _dtype_code(inputs_embeds.dtype), ids.data_ptr(),
                                      _ptr(g), n_geom if g is not None else 0, v.data_ptr(), v.shape[1], B, L, H,
                                      int(image_id), _ptr(overflow), _stream()),
        "vq_splice_features")
    return inputs_embeds


def splice_features_backward(grad: torch.Tensor, input_ids: torch.Tensor, n_geom: int, n_vis: int, image_id: int,
                             need_embeds: bool = True, need_geom: bool = True, need_vis: bool = True):
    """Backward of `splice_features`: (d_embeds [B,L,H] in grad's dtype, d_geom f32 [B,H], d_vis f32 [B,n_vis,H])."""
    B, L, H = grad.shape
    g = grad.contiguous()
    ids = input_ids.to(torch.int64).contiguous()
    d_emb = torch.empty_like(g) if need_embeds else None
    d_geom = torch.empty(B, H, dtype=torch.float32, device=g.device) if (need_geom and n_geom > 0) else None
    d_vis = torch.empty(B, n_vis, H, dtype=torch.float32, device=g.device) if need_vis else None
    _lib.check(
        _lib.lib().vq_splice_features_backward(g.data_ptr(), _dtype_code(g.dtype), ids.data_ptr(), _ptr(d_emb),
                                               _ptr(d_geom), _ptr(d_vis), n_geom, n_vis, B, L, H, int(image_id),
                                               _stream()),
        "vq_splice_features_backward")
    return d_emb, d_geom, d_vis


def geom_head_forward(parts, widths, w1, b1, w2, b2, B: int, V: int):
    """(y, pooled, pre, act) of encode_geom + geom_head (src/models/vggt_qwen3_vlm.py:51-56,164-177); `parts` = four f32
    [B, V, widths[k]] tensors or None (zeros)."""
    H, in_dim = w1.shape
    dev = w1.device
    for t in (w1, b1, w2, b2):
        _req(t, torch.float32, "geom_head parameter")
    keep = [None if p is None else p.to(device=dev, dtype=torch.float32).contiguous() for p in parts]
    ptrs = (C.c_void_p * 4)(*[None if p is None else p.data_ptr() for p in keep])
    wid = (C.c_int32 * 4)(*[int(w) for w in widths])
    pooled = torch.empty(B, in_dim, dtype=torch.float32, device=dev)
    pre = torch.empty(B, H, dtype=torch.float32, device=dev)
    act = torch.empty(B, H, dtype=torch.float32, device=dev)
    y = torch.empty(B, H, dtype=torch.float32, device=dev)
    _lib.check(
        _lib.lib().vq_geom_head_forward(ptrs, wid, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                        pooled.data_ptr(), pre.data_ptr(), act.data_ptr(), y.data_ptr(), B, V, in_dim, H,
                                        _stream()),
        "vq_geom_head_forward")
    return y, pooled, pre, act


def geom_head_backward(dy, pooled, pre, act, w2):
    """(dw1, db1, dw2, db2) f32."""
    B, H = dy.shape
    in_dim = pooled.shape[1]
    dev = dy.device
    dy = dy.to(torch.float32).contiguous()
    dw1 = torch.empty(H, in_dim, dtype=torch.float32, device=dev)
    db1 = torch.empty(H, dtype=torch.float32, device=dev)
    dw2 = torch.empty(H, H, dtype=torch.float32, device=dev)
    db2 = torch.empty(H, dtype=torch.float32, device=dev)
    scratch = torch.empty(B, H, dtype=torch.float32, device=dev)
    _lib.check(
        _lib.lib().vq_geom_head_backward(dy.data_ptr(), pooled.data_ptr(), pre.data_ptr(), act.data_ptr(), w2.data_ptr(),
                                         dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(),
                                         scratch.data_ptr(), B, in_dim, H, 0, _stream()),
        "vq_geom_head_backward")
    return dw1, db1, dw2, db2


def splice_insert(inputs_embeds: torch.Tensor, attention_mask: torch.Tensor, input_ids: torch.Tensor,
                  vis_tokens: torch.Tensor, image_id: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inference splice (src/inference/qa_inference.py:119-145): the FIRST <image> token (in (batch, position) order)
    is replaced in every sample by the vision tokens, the mask is extended with ones; inputs are returned unchanged
    when there is no <image> token.  Two launches; the one 8-byte read-back replaces the reference's `.nonzero()`."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("splice_insert: CUDA tensors only")
    B, L, Hd = inputs_embeds.shape
    n = vis_tokens.shape[1]
    emb = inputs_embeds.contiguous()
    vis = vis_tokens.to(emb.dtype).contiguous()
    mask = attention_mask.contiguous()
    ids = input_ids.to(torch.int64).contiguous()
    first = torch.full((1,), B * L, dtype=torch.int64, device=emb.device)
    out = torch.empty((B, L - 1 + n, Hd), dtype=emb.dtype, device=emb.device)
    out_mask = torch.empty((B, L - 1 + n), dtype=mask.dtype, device=emb.device)
    _lib.check(_lib.lib().vq_find_first_token(ids.data_ptr(), B * L, image_id, first.data_ptr(), _stream()),
               "vq_find_first_token")
    _lib.check(
        _lib.lib().vq_splice_insert(emb.data_ptr(), _dtype_code(emb.dtype), mask.data_ptr(), mask.element_size(),
                                    vis.data_ptr(), first.data_ptr(), out.data_ptr(), out_mask.data_ptr(), B, L, Hd, n,
                                    _stream()),
        "vq_splice_insert")
    if int(first.item()) >= B * L:
        return inputs_embeds, attention_mask
    return out, out_mask


def adamw_step(param: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, grad: torch.Tensor, step: int, *,
               lr: float, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
               grad_scale: float = 1.0, param_bf16: Optional[torch.Tensor] = None) -> None:
    """In-place fused AdamW (torch.optim.AdamW semantics) on flat f32 param / state buffers; `grad` f32 or bf16."""
    for t, name in ((param, "param"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, name)
    n = param.numel()
    assert exp_avg.numel() == n and exp_avg_sq.numel() == n and grad.numel() == n
    assert param.is_contiguous() and exp_avg.is_contiguous() and exp_avg_sq.is_contiguous() and grad.is_contiguous()
    if param_bf16 is not None:
        _req(param_bf16, torch.bfloat16, "param_bf16")
        assert param_bf16.numel() == n and param_bf16.is_contiguous()
    _lib.check(
        _lib.lib().vq_adamw_step(param.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), grad.data_ptr(),
                                 _dtype_code(grad.dtype), param_bf16.data_ptr() if param_bf16 is not None else None, n,
                                 step, lr, betas[0], betas[1], eps, weight_decay, grad_scale, _stream()),
        "vq_adamw_step")


def adamw_reduce_step(master: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, grad_srcs, weight_out, step: int,
                      *, lr: float, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                      grad_scale: float = 1.0) -> None:
    """AdamW on a flat f32 master slice whose gradient is the fp32 SUM of `grad_srcs` (1..16 equally long bf16 or f32
    tensors: this rank's slice + the peers' slices); `weight_out` (bf16 or f32, optional) receives the updated compute
    copy.  The reduce step of a reduce-scatter fused into the optimizer (vq_adamw_reduce_step)."""
    for t, name in ((master, "master"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, name)
    n = master.numel()
    gdt = grad_srcs[0].dtype
    for g in grad_srcs:
        if not g.is_cuda or g.dtype != gdt or g.numel() != n or not g.is_contiguous():
            raise RuntimeError("adamw_reduce_step: gradient sources must be contiguous CUDA tensors of one dtype and length")
    ptrs = (C.c_void_p * len(grad_srcs))(*[g.data_ptr() for g in grad_srcs])
    wdt = VQ_BF16
    if weight_out is not None:
        assert weight_out.numel() == n and weight_out.is_contiguous()
        wdt = _dtype_code(weight_out.dtype)
    _lib.check(
        _lib.lib().vq_adamw_reduce_step(master.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), ptrs,
                                        len(grad_srcs), _dtype_code(gdt), _ptr(weight_out), wdt, n, int(step), lr,
                                        betas[0], betas[1], eps, weight_decay, grad_scale, _stream()),
        "vq_adamw_reduce_step")


def cast_f32_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst (bf16) = src (f32), flat contiguous CUDA tensors of equal length."""
    _req(src, torch.float32, "src")
    _req(dst, torch.bfloat16, "dst")
    assert src.numel() == dst.numel() and src.is_contiguous() and dst.is_contiguous()
    _lib.check(_lib.lib().vq_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vq_cast_f32_bf16")


def preprocess_views(views_u8: torch.Tensor, image_size: int) -> torch.Tensor:
    """`build_default_transform(image_size)` (src/dataio/collate_multiview.py:12-19) on a batch of equally sized RGB
    views: uint8 [..., H, W, 3] on the GPU -> float32 [..., 3, image_size, image_size] in [0, 1], bit-exact with
    torchvision-on-PIL (antialiased bicubic resize of the shorter edge, centre crop, ToTensor)."""
    if not views_u8.is_cuda or views_u8.dtype != torch.uint8:
        raise RuntimeError("preprocess_views: CUDA uint8 tensors only")
    if views_u8.dim() < 3 or views_u8.shape[-1] != 3:
        raise RuntimeError("preprocess_views: expected [..., H, W, 3]")
    lead = tuple(views_u8.shape[:-3])
    H, W = int(views_u8.shape[-3]), int(views_u8.shape[-2])
    x = views_u8.reshape(-1, H, W, 3).contiguous()
    frames = x.shape[0]
    nbytes = _lib.lib().vq_preprocess_workspace_bytes(frames, H, W, image_size)
    if nbytes == 0:
        raise RuntimeError(_lib.lib().vq_last_error().decode())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty((frames, 3, image_size, image_size), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().vq_preprocess_views(x.data_ptr(), out.data_ptr(), frames, H, W, image_size, ws.data_ptr(),
                                              nbytes, _stream()), "vq_preprocess_views")
    return out.reshape(lead + (3, image_size, image_size))
