"""Host-side mirror of the VGGT `Aggregator` module, executing on libvq_b200.so.

Interface kept from the reference call site (src/models/vggt_qwen3_vlm.py:72-83,144-156):
  * `VGGT(img_size=518, patch_size=14, embed_dim=1024, ...)` exposes `.aggregator(images) -> (list, patch_start_idx)`
    with `list[-1]` of shape [B, S, P, 2*embed_dim];
  * state-dict keys are the upstream ones, so `vggt_1B_commercial.pt` loads unchanged
    (`strict=False`; the camera/point/depth/track heads of the full VGGT are never run on this path and are ignored).

The modules below are parameter containers only: no PyTorch compute runs in `forward`; the whole forward is one
`vq_aggregator_forward` call (csrc/aggregator.cu).  There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import AggregatorCfg, AggregatorWeights, BlockWeights


class _LayerScale(nn.Module):
    def __init__(self, dim: int, init_values: float) -> None:
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class _AttnParams(nn.Module):
    def __init__(self, dim: int, head_dim: int, qk_norm: bool) -> None:
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)
        if qk_norm:
            self.q_norm = nn.LayerNorm(head_dim)
            self.k_norm = nn.LayerNorm(head_dim)


class _MlpParams(nn.Module):
    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)


class _BlockParams(nn.Module):
    """Parameters of one pre-norm block ([UPSTREAM] vggt/layers/block.py `Block`)."""

    def __init__(self, dim: int, heads: int, hidden: int, init_values: float, qk_norm: bool, eps: float) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _AttnParams(dim, dim // heads, qk_norm)
        self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _MlpParams(dim, hidden)
        self.ls2 = _LayerScale(dim, init_values)


class _PatchProj(nn.Module):
    def __init__(self, dim: int, patch: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)  # container for weight [D,3,p,p] + bias


class _DinoStem(nn.Module):
    """Parameters of the DINOv2 ViT-with-registers stem ([UPSTREAM] vggt/layers/vision_transformer.py)."""

    def __init__(self, img_size: int, patch: int, dim: int, depth: int, heads: int, hidden: int, registers: int) -> None:
        super().__init__()
        g = img_size // patch
        self.patch_embed = _PatchProj(dim, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, g * g + 1, dim))
        self.register_tokens = nn.Parameter(torch.zeros(1, registers, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, dim))
        self.blocks = nn.ModuleList(
            [_BlockParams(dim, heads, hidden, 1.0, qk_norm=False, eps=1e-6) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        nn.init.normal_(self.register_tokens, std=1e-6)
        for m in self.blocks.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


class Aggregator(nn.Module):
    """Drop-in for `vggt.models.aggregator.Aggregator` (forward only, frozen)."""

    def __init__(self, img_size: int = 518, patch_size: int = 14, embed_dim: int = 1024, depth: int = 24,
                 num_heads: int = 16, mlp_ratio: float = 4.0, num_register_tokens: int = 4,
                 rope_freq: float = 100.0, init_values: float = 0.01, stem_depth: Optional[int] = None) -> None:
        super().__init__()
        if embed_dim % num_heads or embed_dim // num_heads != 64:
            raise ValueError("the sm_100a attention kernel is specialised for head_dim 64")
        self.img_size, self.patch_size, self.embed_dim = img_size, patch_size, embed_dim
        self.depth, self.num_heads = depth, num_heads
        self.hidden = int(embed_dim * mlp_ratio)
        self.num_register_tokens = num_register_tokens
        self.patch_start_idx = 1 + num_register_tokens
        self.rope_freq = rope_freq
        self.aa_order = ["frame", "global"]
        stem_depth = depth if stem_depth is None else stem_depth
        self.patch_embed = _DinoStem(img_size, patch_size, embed_dim, stem_depth, num_heads, self.hidden,
                                     num_register_tokens)
        self.frame_blocks = nn.ModuleList(
            [_BlockParams(embed_dim, num_heads, self.hidden, init_values, True, 1e-5) for _ in range(depth)])
        self.global_blocks = nn.ModuleList(
            [_BlockParams(embed_dim, num_heads, self.hidden, init_values, True, 1e-5) for _ in range(depth)])
        self.camera_token = nn.Parameter(torch.randn(1, 2, 1, embed_dim) * 1e-6)
        self.register_token = nn.Parameter(torch.randn(1, 2, num_register_tokens, embed_dim) * 1e-6)
        self.Kp = (3 * patch_size * patch_size + 7) // 8 * 8
        # Optional: LayerNorm folded into the qkv / fc1 GEMMs (no standalone LayerNorm launches; csrc/aggregator.cu).
        # OFF by default: measured on B200 (round 2) the K = 1024 GEMMs are epilogue-issue-bound, so the folded arithmetic
        # costs more there (+47 us per block at c2) than the two LayerNorm launches it removes (33 us), and the f32
        # atomics of the row statistics make the result run-to-run non-bit-identical.  VQ_FOLD_LN=1 or
        # `fold_layernorm = True` + invalidate_packed() selects it (parity-tested both ways).
        self.fold_layernorm = os.environ.get("VQ_FOLD_LN", "0") == "1"
        # Default schedule: residual adds deferred into the LayerNorm pass after them (vq_residual_layernorm); the proj /
        # fc2 GEMMs then write a bf16 branch output with a plain epilogue.  VQ_DEFERRED_RESIDUAL=0 = round-1 schedule
        # (f32 read-modify-write in the GEMM epilogue + standalone LayerNorm).
        self.deferred_residual = os.environ.get("VQ_DEFERRED_RESIDUAL", "1") != "0"
        self._packed: Optional[dict] = None
        self._workspaces: Dict[tuple, torch.Tensor] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # ------------------------------------------------------------------ weight packing
    def invalidate_packed(self) -> None:
        self._packed = None

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .bfloat16() move the masters: repack lazily
        self._packed = None
        self._workspaces = {}
        return super()._apply(fn, *args, **kwargs)

    def _pack(self, device: torch.device, gh: int, gw: int) -> dict:
        """bf16 copies of the matrices (nn.Linear layout), f32 copies of every vector, rope tables, and the
        ctypes structs pointing at them.  Done once per (device, patch grid); weights are frozen."""
        key = (str(device), gh, gw, self.fold_layernorm)
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        keep: List[torch.Tensor] = []

        def mat(t: torch.Tensor) -> int:
            t = t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
            keep.append(t)
            return t.data_ptr()

        def vec(t: torch.Tensor) -> int:
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        def folded(lin: nn.Linear, ln: nn.LayerNorm):
            """LayerNorm folded into the Linear that consumes it (vq_gemm_desc.ln_stats): W' = bf16(W diag(gamma)),
            colsum = row sums of the ROUNDED W' in f32 (what the tensor core multiplies), b' = W beta + b."""
            wf = lin.weight.detach().to(device=device, dtype=torch.float32)
            gam = ln.weight.detach().to(device=device, dtype=torch.float32)
            bet = ln.bias.detach().to(device=device, dtype=torch.float32)
            w16 = (wf * gam[None, :]).to(torch.bfloat16).contiguous()
            keep.append(w16)
            colsum = w16.float().sum(dim=1)
            bias = wf @ bet + lin.bias.detach().to(device=device, dtype=torch.float32)
            return w16.data_ptr(), vec(bias), vec(colsum)

        def block(b: _BlockParams) -> BlockWeights:
            w = BlockWeights()
            w.norm1_w, w.norm1_b = vec(b.norm1.weight), vec(b.norm1.bias)
            if self.fold_layernorm:
                w.ln_folded = 1
                w.qkv_w, w.qkv_b, w.qkv_colsum = folded(b.attn.qkv, b.norm1)
            else:
                w.qkv_w, w.qkv_b = mat(b.attn.qkv.weight), vec(b.attn.qkv.bias)
            if hasattr(b.attn, "q_norm"):
                w.q_norm_w, w.q_norm_b = vec(b.attn.q_norm.weight), vec(b.attn.q_norm.bias)
                w.k_norm_w, w.k_norm_b = vec(b.attn.k_norm.weight), vec(b.attn.k_norm.bias)
                w.qk_score_bound = _qk_score_bound(b.attn.q_norm, b.attn.k_norm)
            w.proj_w, w.proj_b = mat(b.attn.proj.weight), vec(b.attn.proj.bias)
            w.ls1 = vec(b.ls1.gamma)
            w.norm2_w, w.norm2_b = vec(b.norm2.weight), vec(b.norm2.bias)
            if self.fold_layernorm:
                w.fc1_w, w.fc1_b, w.fc1_colsum = folded(b.mlp.fc1, b.norm2)
            else:
                w.fc1_w, w.fc1_b = mat(b.mlp.fc1.weight), vec(b.mlp.fc1.bias)
            w.fc2_w, w.fc2_b = mat(b.mlp.fc2.weight), vec(b.mlp.fc2.bias)
            w.ls2 = vec(b.ls2.gamma)
            return w

        stem = self.patch_embed
        D = self.embed_dim
        aw = AggregatorWeights()
        conv_w = stem.patch_embed.proj.weight.detach().float().reshape(D, -1)
        aw.patch_w = mat(F.pad(conv_w, (0, self.Kp - conv_w.shape[1])))
        aw.patch_b = vec(stem.patch_embed.proj.bias)
        aw.cls_token = vec(stem.cls_token.reshape(D))
        aw.pos_embed = vec(self._interpolated_pos_embed(gh, gw).reshape(1 + gh * gw, D))
        aw.register_tokens = vec(stem.register_tokens.reshape(-1, D))
        stem_arr = (BlockWeights * len(stem.blocks))(*[block(b) for b in stem.blocks])
        frame_arr = (BlockWeights * self.depth)(*[block(b) for b in self.frame_blocks])
        global_arr = (BlockWeights * self.depth)(*[block(b) for b in self.global_blocks])
        aw.stem_blocks, aw.n_stem = stem_arr, len(stem.blocks)
        aw.stem_norm_w, aw.stem_norm_b = vec(stem.norm.weight), vec(stem.norm.bias)
        aw.camera_token = vec(self.camera_token.reshape(2, 1, D))
        aw.register_token = vec(self.register_token.reshape(2, -1, D))
        aw.frame_blocks, aw.global_blocks, aw.depth = frame_arr, global_arr, self.depth
        max_pos = max(gh, gw)
        cos, sin = _rope_tables(max_pos, 64, self.rope_freq)
        aw.rope_cos, aw.rope_sin, aw.rope_max_pos = vec(cos), vec(sin), max_pos
        self._packed = {"key": key, "weights": aw, "keep": keep, "arrays": (stem_arr, frame_arr, global_arr)}
        return self._packed

    def _interpolated_pos_embed(self, gh: int, gw: int) -> torch.Tensor:
        """DINOv2 interpolate_pos_encoding (antialias bicubic, offset 0) evaluated once per patch grid on the
        frozen parameter — host-side parameter preparation, not part of the per-step path."""
        pe = self.patch_embed.pos_embed.detach().float()
        n = pe.shape[1] - 1
        m = int(round(math.sqrt(n)))
        if gh * gw == n and gh == gw:
            return pe
        patch = pe[:, 1:].reshape(1, m, m, -1).permute(0, 3, 1, 2)
        patch = F.interpolate(patch.cpu(), size=(gh, gw), mode="bicubic", antialias=True)
        patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, -1).to(pe.device)
        return torch.cat([pe[:, :1], patch], dim=1)

    # ------------------------------------------------------------------ execution
    def _cfg(self, B: int, S: int, H: int, W: int) -> AggregatorCfg:
        c = AggregatorCfg()
        c.B, c.S, c.H, c.W, c.patch = B, S, H, W, self.patch_size
        c.D, c.heads, c.hidden = self.embed_dim, self.num_heads, self.hidden
        c.n_special, c.Kp = self.patch_start_idx, self.Kp
        c.stem_eps, c.block_eps, c.qk_eps = 1e-6, 1e-5, 1e-5
        c.deferred_residual = 1 if self.deferred_residual else 0
        return c

    def _workspace(self, cfg: AggregatorCfg, device: torch.device) -> torch.Tensor:
        key = (str(device), cfg.B, cfg.S, cfg.H, cfg.W)
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.lib().vq_aggregator_workspace_bytes(C.byref(cfg))
            if nbytes == 0:
                _lib.check(-1, "vq_aggregator_workspace_bytes")
            self._workspaces = {key: torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)}
            ws = self._workspaces[key]
        return ws

    @torch.no_grad()
    def encode_tokens(self, images: torch.Tensor, n_take: Optional[int] = None, return_all: bool = False):
        """images [B,S,3,H,W] in [0,1] -> bf16 [B, n_take, 2D]: the first n_take tokens of the view-flattened last
        aggregator output (what encode_images keeps).  With return_all, also the list of every iteration's
        [B,S,P,2D] output."""
        u8 = images.dtype == torch.uint8
        if images.dim() != 5 or images.shape[4 if u8 else 2] != 3:
            raise ValueError(f"expected float images [B,S,3,H,W] or uint8 views [B,S,H,W,3], got {tuple(images.shape)}")
        if not images.is_cuda:
            raise RuntimeError("Aggregator runs on sm_100a only: images must be a CUDA tensor (no CPU fallback)")
        if u8:      # views as decoded (HWC bytes): x / 255 happens inside the patch kernel, no f32 image is built
            B, S, H, W, _ = images.shape
        else:
            B, S, _, H, W = images.shape
        ps = self.patch_size
        if H % ps or W % ps:
            raise ValueError(f"image size {(H, W)} must be a multiple of the patch size {ps}")
        gh, gw = H // ps, W // ps
        P = self.patch_start_idx + gh * gw
        n_take = S * P if n_take is None else min(int(n_take), S * P)
        dev = images.device
        img = images.detach().contiguous() if u8 else images.detach().to(torch.float32).contiguous()
        packed = self._pack(dev, gh, gw)
        cfg = self._cfg(B, S, H, W)
        cfg.images_u8 = 1 if u8 else 0
        ws = self._workspace(cfg, dev)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        out = torch.empty(B, n_take, 2 * self.embed_dim, dtype=torch.bfloat16, device=dev)
        all_out: Optional[List[torch.Tensor]] = None
        all_ptrs = None
        if return_all:
            all_out = [torch.empty(B, S * P, 2 * self.embed_dim, dtype=torch.bfloat16, device=dev)
                       for _ in range(self.depth)]
            all_ptrs = (C.c_void_p * self.depth)(*[t.data_ptr() for t in all_out])
        rc = _lib.lib().vq_aggregator_forward(
            C.byref(cfg), C.byref(packed["weights"]), img.data_ptr(), out.data_ptr(), n_take, all_ptrs,
            base, ws.numel() - (base - ws.data_ptr()), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "vq_aggregator_forward")
        if return_all:
            return out, [t.view(B, S, P, -1) for t in all_out]
        return out

    def forward(self, images: torch.Tensor, return_all: bool = False) -> Tuple[List[torch.Tensor], int]:
        """Upstream signature: (aggregated_tokens_list, patch_start_idx).  By default the list holds only the last
        iteration's [B,S,P,2D] tensor — the only entry the reference reads (src/models/vggt_qwen3_vlm.py:148) —
        so 23 dead 5.6 MB/frame concats are skipped; `return_all=True` materialises all `depth` entries."""
        B, S = images.shape[:2]   # float [B,S,3,H,W] or uint8 [B,S,H,W,3]
        if return_all:
            _, outs = self.encode_tokens(images, None, return_all=True)
            return outs, self.patch_start_idx
        last = self.encode_tokens(images, None)
        return [last.view(B, S, -1, last.shape[-1])], self.patch_start_idx


def _qk_score_bound(q_norm: nn.LayerNorm, k_norm: nn.LayerNorm) -> float:
    """Proven upper bound of |q . k| * log2(e)/sqrt(64) for per-head LayerNorm-ed q and k: a normalised 64-vector has
    norm <= 8, so |LN(x)| <= 8 max|w| + |b|_2; RoPE is a rotation and keeps norms.  The attention kernel uses it to
    skip the softmax's range check (vq_attention_fwd_bounded); a margin of 2 % covers the bf16 rounding of q and k."""
    def norm_bound(ln: nn.LayerNorm) -> float:
        w = ln.weight.detach().float()
        b = ln.bias.detach().float() if ln.bias is not None else torch.zeros_like(w)
        return float(math.sqrt(w.numel()) * w.abs().max() + b.norm())

    return 1.02 * norm_bound(q_norm) * norm_bound(k_norm) * math.log2(math.e) / 8.0


def _rope_tables(max_pos: int, head_dim: int, freq: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 cos/sin [max_pos+1, head_dim/4] of the 2-D RoPE ([UPSTREAM] vggt/layers/rope.py, frequency 100):
    inv_freq_j = freq^(-2j/(head_dim/2)); position 0 (special tokens) is the identity rotation."""
    half = head_dim // 2
    j = torch.arange(0, half, 2, dtype=torch.float64)
    inv = 1.0 / (freq ** (j / half))
    ang = torch.arange(max_pos + 1, dtype=torch.float64)[:, None] * inv[None, :]
    return ang.cos().float(), ang.sin().float()


class VGGT(nn.Module):
    """The slice of `vggt.models.vggt.VGGT` this path touches: `.aggregator` and `.embed_dim`
    (src/models/vggt_qwen3_vlm.py:75-83,109).  The prediction heads are not built: the reference never calls
    them here (`:142-144`), and their checkpoint keys are skipped by `strict=False` loading."""

    def __init__(self, img_size: int = 518, patch_size: int = 14, embed_dim: int = 1024, enable_camera: bool = True,
                 enable_point: bool = True, enable_depth: bool = True, enable_track: bool = True, **agg_kwargs) -> None:
        super().__init__()
        self.aggregator = Aggregator(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, **agg_kwargs)
        self.embed_dim = 2 * embed_dim
