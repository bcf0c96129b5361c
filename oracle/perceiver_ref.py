"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference Perceiver projector,
`encode_images` and the inputs_embeds splice.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, in plain matmuls on a state dict that uses the reference's own key names:
  * PerceiverProjector.forward      src/models/projector_perceiver.py:70-82
  * PerceiverLayer.forward          src/models/projector_perceiver.py:44-50
      attn_out,_ = MHA(latents, context, context)   (nn.MultiheadAttention, batch_first, packed in_proj [3E,E]:
      q/k/v = rows [0:E],[E:2E],[2E:3E]; heads = contiguous E/heads chunks; scale 1/sqrt(E/heads))
      latents = LN1(latents + attn_out); latents = LN2(latents + W2 gelu_erf(W1 latents))      (LN eps 1e-5)
    train mode: the four dropout sites (:33 attention probabilities, :46, :37, :49) are applied with the SAME
    counter-based masks the CUDA kernels regenerate (`dropout_keep_mask`, a numpy restatement of Philox4x32-10 as
    laid out in csrc/philox.cuh), so train-mode forward and backward are compared value for value, not statistically.
  * _init_weights / parameter shapes src/models/projector_perceiver.py:13-17,56-68
  * VGGTQwen3VLM.encode_images      src/models/vggt_qwen3_vlm.py:128-162  (slice of the last aggregator output)
  * VGGTQwen3VLM.forward splice     src/models/vggt_qwen3_vlm.py:187-195  (overwrite)
  * _insert_vision_tokens           src/inference/qa_inference.py:119-145 (insert)
  * encode_geom / geom_head         src/models/vggt_qwen3_vlm.py:51-56,164-177

PARITY PINNING: pinned.  tests/golden/perceiver_*.npz hold inputs, weights, outputs and gradients produced by
importing the reference's own `PerceiverProjector` from /root/reference (tests/golden/make_golden.py);
tests/test_oracle_cpu.py checks this restatement against them.
Backward is obtained by autograd through this restatement (fp32 or fp64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


@dataclass
class PerceiverShape:
    """Field names follow the reference PerceiverConfig (src/models/projector_perceiver.py:20-27)."""

    latent_dim: int = 4096
    num_latents: int = 128
    num_heads: int = 8
    num_layers: int = 6
    ffn_dim: int = 16384
    dropout: float = 0.1
    in_dim: int = 2048
    out_dim: int = 2560


def _xavier_uniform(gen: torch.Generator, out_f: int, in_f: int) -> torch.Tensor:
    bound = math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound


def init_perceiver_state(shape: PerceiverShape, seed: int = 0, nonzero_bias: bool = False) -> State:
    """Random init with the distributions of the reference (`_init_weights`: xavier-uniform Linear weights, zero
    biases; latents = randn * 0.02; MHA in_proj xavier-uniform over the packed [3E,E] matrix).  The draw order
    differs from torch's module construction, so values are not bit-identical to `PerceiverProjector(seed)`;
    goldens carry the reference's actual weights."""
    gen = torch.Generator().manual_seed(seed)
    E, Fd = shape.latent_dim, shape.ffn_dim
    s: State = {"latents": torch.randn(shape.num_latents, E, generator=gen) * 0.02}

    def lin(name: str, o: int, i: int) -> None:
        s[f"{name}.weight"] = _xavier_uniform(gen, o, i)
        s[f"{name}.bias"] = 0.02 * torch.randn(o, generator=gen) if nonzero_bias else torch.zeros(o)

    lin("in_proj", E, shape.in_dim)
    for l in range(shape.num_layers):
        p = f"layers.{l}"
        s[f"{p}.self_attn.in_proj_weight"] = _xavier_uniform(gen, 3 * E, E)
        s[f"{p}.self_attn.in_proj_bias"] = 0.02 * torch.randn(3 * E, generator=gen) if nonzero_bias else torch.zeros(3 * E)
        lin(f"{p}.self_attn.out_proj", E, E)
        lin(f"{p}.mlp.0", Fd, E)
        lin(f"{p}.mlp.3", E, Fd)
        for n in ("norm1", "norm2"):
            s[f"{p}.{n}.weight"] = 1.0 + (0.1 * torch.randn(E, generator=gen) if nonzero_bias else torch.zeros(E))
            s[f"{p}.{n}.bias"] = 0.05 * torch.randn(E, generator=gen) if nonzero_bias else torch.zeros(E)
    lin("out_proj", shape.out_dim, E)
    return s


# --------------------------------------------------------------------------------------------------
# dropout masks: Philox4x32-10, 8 elements per call (16-bit lanes), as in vggt_qwen3_b200/csrc/philox.cuh
# --------------------------------------------------------------------------------------------------

def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    import numpy as np

    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85, np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0 & MASK, p1 & MASK, n2 & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def dropout_threshold(p: float) -> int:
    return max(1, min(65535, int(p * 65536.0 + 0.5)))


def dropout_keep_mask(rows: int, cols: int, ld: int, p: float, seed: int, offset: int, site: int) -> torch.Tensor:
    """bool [rows, cols]: element (r, c) of a [rows, ld] dropout site is kept iff the 16-bit lane (c % 8) of
    philox(ctr = {group lo, group hi, site, offset lo}, key = {seed lo ^ offset hi, seed hi}), group = r * ld/8 + c/8,
    is >= round(p * 65536)."""
    import numpy as np

    if p <= 0.0:
        return torch.ones(rows, cols, dtype=torch.bool)
    assert ld % 8 == 0 and cols <= ld
    r = np.arange(rows, dtype=np.uint64)[:, None]
    g = r * np.uint64(ld // 8) + np.arange((cols + 7) // 8, dtype=np.uint64)[None, :]
    k0 = (seed & 0xFFFFFFFF) ^ ((offset >> 32) & 0xFFFFFFFF)
    k1 = (seed >> 32) & 0xFFFFFFFF
    out = _philox4x32_10(g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.uint64(site), np.uint64(offset & 0xFFFFFFFF),
                         k0, k1)
    lanes = np.stack([(out[j >> 1] >> np.uint64(16 * (j & 1))) & np.uint64(0xFFFF) for j in range(8)], axis=-1)
    keep = (lanes >= np.uint64(dropout_threshold(p))).reshape(rows, -1)[:, :cols]
    return torch.from_numpy(np.ascontiguousarray(keep))


def _drop(x: torch.Tensor, rows: int, cols: int, ld: int, dropout, site: int) -> torch.Tensor:
    p, seed, offset = dropout
    keep = dropout_keep_mask(rows, cols, ld, p, seed, offset, site).to(x.device)
    return x * keep.reshape(x.shape).to(x.dtype) * (1.0 / (1.0 - p))


def num_layers_of(s: State) -> int:
    return 1 + max(int(k.split(".")[1]) for k in s if k.startswith("layers."))


def perceiver_layer(s: State, p: str, latents: torch.Tensor, context: torch.Tensor, heads: int,
                    dropout=None, layer: int = 0) -> torch.Tensor:
    """PerceiverLayer.forward (src/models/projector_perceiver.py:44-50).  `dropout` = None (eval) or
    (p, seed, offset): train mode with the mask streams 4*layer + {0 attention probabilities, 1 after the attention
    output projection, 2 after GELU, 3 after the MLP}."""
    B, Lq, E = latents.shape
    T = context.shape[1]
    hd = E // heads
    W, bqkv = s[f"{p}.self_attn.in_proj_weight"], s[f"{p}.self_attn.in_proj_bias"]
    q = latents @ W[:E].t() + bqkv[:E]
    k = context @ W[E:2 * E].t() + bqkv[E:2 * E]
    v = context @ W[2 * E:].t() + bqkv[2 * E:]
    q = q.reshape(B, Lq, heads, hd).transpose(1, 2)
    k = k.reshape(B, T, heads, hd).transpose(1, 2)
    v = v.reshape(B, T, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    drop = dropout is not None and dropout[0] > 0.0
    if drop:   # rows = (sample, head, latent), row length padded to a multiple of 64 keys like the kernel's buffer
        att = _drop(att, B * heads * Lq, T, (T + 63) // 64 * 64, dropout, 4 * layer + 0)
    a = (att @ v).transpose(1, 2).reshape(B, Lq, E)
    a = a @ s[f"{p}.self_attn.out_proj.weight"].t() + s[f"{p}.self_attn.out_proj.bias"]
    if drop:
        a = _drop(a, B * Lq, E, E, dropout, 4 * layer + 1)
    x = F.layer_norm(latents + a, (E,), s[f"{p}.norm1.weight"], s[f"{p}.norm1.bias"], 1e-5)
    h = F.gelu(x @ s[f"{p}.mlp.0.weight"].t() + s[f"{p}.mlp.0.bias"])
    if drop:
        h = _drop(h, B * Lq, h.shape[-1], h.shape[-1], dropout, 4 * layer + 2)
    h = h @ s[f"{p}.mlp.3.weight"].t() + s[f"{p}.mlp.3.bias"]
    if drop:
        h = _drop(h, B * Lq, E, E, dropout, 4 * layer + 3)
    return F.layer_norm(x + h, (E,), s[f"{p}.norm2.weight"], s[f"{p}.norm2.bias"], 1e-5)


def perceiver_forward(s: State, tokens: torch.Tensor, heads: int, dropout=None) -> torch.Tensor:
    """PerceiverProjector.forward (src/models/projector_perceiver.py:70-82): tokens [B,T,in_dim] -> [B,num_latents,out_dim].
    `dropout` = (p, seed, offset) selects train mode with the kernels' masks."""
    B = tokens.shape[0]
    context = tokens @ s["in_proj.weight"].t() + s["in_proj.bias"]
    latents = s["latents"].unsqueeze(0).expand(B, -1, -1)
    for l in range(num_layers_of(s)):
        latents = perceiver_layer(s, f"layers.{l}", latents, context, heads, dropout, l)
    return latents @ s["out_proj.weight"].t() + s["out_proj.bias"]


def perceiver_forward_backward(s: State, tokens: torch.Tensor, grad_out: torch.Tensor, heads: int,
                               dtype: torch.dtype = torch.float32, dropout=None,
                               ) -> Tuple[torch.Tensor, State, torch.Tensor]:
    """Forward + autograd backward of the restatement. Returns (out, grads by key, grad wrt tokens)."""
    leaf = {k: v.detach().to(dtype).requires_grad_(True) for k, v in s.items()}
    tok = tokens.detach().to(dtype).requires_grad_(True)
    out = perceiver_forward(leaf, tok, heads, dropout)
    out.backward(grad_out.to(dtype))
    return out.detach(), {k: v.grad for k, v in leaf.items()}, tok.grad


def perceiver_flops(shape: PerceiverShape, T: int) -> float:
    """Forward FLOPs per sample: 315.68 GF for perceiver_small at T = 128."""
    E, Fd, Lq = shape.latent_dim, shape.ffn_dim, shape.num_latents
    hd = E // shape.num_heads
    per_layer = 2.0 * Lq * E * E + 2.0 * T * E * 2 * E + 2.0 * Lq * E * E + 4.0 * Lq * T * hd * shape.num_heads + 4.0 * Lq * E * Fd
    return 2.0 * T * shape.in_dim * E + shape.num_layers * per_layer + 2.0 * Lq * E * shape.out_dim


# --------------------------------------------------------------------------------------------------
# encode_images slice, splice, geom head
# --------------------------------------------------------------------------------------------------

def select_vis_tokens(agg_last: torch.Tensor, num_vis_tokens: int) -> torch.Tensor:
    """src/models/vggt_qwen3_vlm.py:148-156: rank-3 -> first tokens; rank-4 [B,V,P,C] -> flatten views, first tokens."""
    if agg_last.dim() == 3:
        return agg_last[:, :num_vis_tokens, :]
    B = agg_last.shape[0]
    return agg_last.reshape(B, -1, agg_last.shape[-1])[:, :num_vis_tokens, :]


def splice_overwrite(inputs_embeds: torch.Tensor, input_ids: torch.Tensor, features: torch.Tensor, image_id: int) -> torch.Tensor:
    """Training splice (src/models/vggt_qwen3_vlm.py:190-195): for every <image> position, in (batch, pos) order,
    overwrite inputs_embeds[b, pos:pos+n] with features[b] (python slice semantics: clipped at the sequence end)."""
    out = inputs_embeds.clone()
    n = features.shape[1]
    for b, pos in (input_ids == image_id).nonzero(as_tuple=False).tolist():
        room = out.shape[1] - pos
        if room < n:
            # the reference's slice assignment raises a shape-mismatch RuntimeError here
            raise RuntimeError(f"splice: span of {n} tokens does not fit at position {pos} of {out.shape[1]}")
        out[b, pos:pos + n] = features[b].to(out.dtype)
    return out


def splice_insert(inputs_embeds: torch.Tensor, attention_mask: torch.Tensor, input_ids: torch.Tensor,
                  vis: torch.Tensor, image_id: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inference splice (src/inference/qa_inference.py:119-145), B = 1: the first <image> token is replaced by the
    vision tokens (sequence grows by n-1), mask extended with ones; unchanged if no <image> token."""
    pos = (input_ids[0] == image_id).nonzero(as_tuple=False)
    if pos.numel() == 0:
        return inputs_embeds, attention_mask
    p = int(pos[0])
    emb = torch.cat([inputs_embeds[:, :p], vis.to(inputs_embeds.dtype), inputs_embeds[:, p + 1:]], dim=1)
    ones = torch.ones(attention_mask.shape[0], vis.shape[1], dtype=attention_mask.dtype)
    mask = torch.cat([attention_mask[:, :p], ones, attention_mask[:, p + 1:]], dim=1)
    return emb, mask


def encode_geom(s: State, geom_token: Optional[Dict[str, torch.Tensor]], geom_tokens: int) -> Optional[torch.Tensor]:
    """src/models/vggt_qwen3_vlm.py:164-177 with geom_head = Linear(37,H) - SiLU - Linear(H,H) (keys geom_head.{0,2}.*)."""
    if not geom_token or geom_tokens == 0:
        return None
    feats = torch.cat([geom_token["R"], geom_token["t"], geom_token["K"], geom_token["depth_hist"]], dim=-1)
    h = feats.mean(dim=1) @ s["geom_head.0.weight"].t() + s["geom_head.0.bias"]
    h = F.silu(h) @ s["geom_head.2.weight"].t() + s["geom_head.2.bias"]
    return h.unsqueeze(1).expand(-1, geom_tokens, -1)
