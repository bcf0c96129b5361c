"""ORACLE (test infrastructure, not product code) — fp32 CPU restatement of the VGGT Aggregator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (vggt_qwen3_b200/) never does.

What it restates
----------------
`VGGT(...).aggregator(images)` as called by the reference at src/models/vggt_qwen3_vlm.py:144, constructed
with img_size=518, patch_size=14, embed_dim=1024 (src/models/vggt_qwen3_vlm.py:75-83).  The `vggt` package
(facebookresearch/vggt, fork Sycamorers/vggt; third_party/README.md:10) is NOT vendored in the reference
repo, not pinned to any commit (env/environment.yml:19-32 does not list it) and not installed here, so this
file restates its published algorithm (vggt/models/aggregator.py, vggt/layers/{block,attention,rope,
layer_scale,mlp,patch_embed,vision_transformer}.py) from SURVEY.md Appendix A:

  A.2  normalise -> DINOv2 ViT-L/14 with 4 registers (24 pre-norm blocks, LN eps 1e-6, LayerScale 1.0)
       -> x_norm_patchtokens -> prepend camera/register tokens -> 24 x (frame block, global block)
  A.3  aggregator block: pre-norm (eps 1e-5), fused qkv, per-head LayerNorm(64) on q and k, 2-D RoPE,
       SDPA (scale 1/8), proj, LayerScale(0.01), MLP fc1-GELU(erf)-fc2, LayerScale
  A.4  2-D RoPE, frequency 100, fp32 tables (upstream builds them in the activation dtype)
  A.6  state-dict key names (used verbatim so real `vggt_1B_commercial.pt` weights would load)

PARITY PINNING: the aggregator as a whole is **parity unpinned** — the reference holds no test, golden
vector or fixture for it and the upstream package cannot be run here.  The DINOv2 stem
(`dinov2_forward`) IS pinned: tests/test_oracle_cpu.py checks it against HuggingFace
`Dinov2WithRegistersModel` (same upstream lineage, present in this image) with mapped weights.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

_RESNET_MEAN = (0.485, 0.456, 0.406)
_RESNET_STD = (0.229, 0.224, 0.225)


@dataclass
class AggregatorConfig:
    """Upstream `Aggregator.__init__` arguments."""

    img_size: int = 518
    patch_size: int = 14
    embed_dim: int = 1024
    depth: int = 24            # aggregator iterations: `depth` frame blocks + `depth` global blocks
    stem_depth: int = 24       # DINOv2 ViT-L blocks
    num_heads: int = 16
    mlp_ratio: float = 4.0
    num_register_tokens: int = 4
    rope_freq: float = 100.0
    init_values: float = 0.01  # aggregator LayerScale init
    stem_init_values: float = 1.0
    aa_order: Tuple[str, ...] = ("frame", "global")

    @property
    def patch_start_idx(self) -> int:
        return 1 + self.num_register_tokens

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def hidden_dim(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)


# --------------------------------------------------------------------------------------------------
# parameter construction (random init of the named architecture, upstream key names)
# --------------------------------------------------------------------------------------------------

def _linear_default(gen: torch.Generator, out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.Linear default init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def _trunc_normal(gen: torch.Generator, shape, std: float) -> torch.Tensor:
    t = torch.randn(*shape, generator=gen) * std
    return t.clamp_(-2.0, 2.0)  # timm trunc_normal_(std=.02) truncates at +-2 absolute: never hit at std .02


def _block_params(state: State, prefix: str, cfg: AggregatorConfig, gen: torch.Generator, *, vit_init: bool,
                  qk_norm: bool, ls_init: float) -> None:
    D, Hd = cfg.embed_dim, cfg.hidden_dim
    for name, (o, i) in {"attn.qkv": (3 * D, D), "attn.proj": (D, D), "mlp.fc1": (Hd, D), "mlp.fc2": (D, Hd)}.items():
        if vit_init:  # DINOv2 init_weights_vit_timm: trunc_normal(std=.02), zero bias
            w, b = _trunc_normal(gen, (o, i), 0.02), torch.zeros(o)
        else:
            w, b = _linear_default(gen, o, i)
        state[f"{prefix}.{name}.weight"], state[f"{prefix}.{name}.bias"] = w, b
    for n in ("norm1", "norm2"):
        state[f"{prefix}.{n}.weight"], state[f"{prefix}.{n}.bias"] = torch.ones(D), torch.zeros(D)
    state[f"{prefix}.ls1.gamma"] = torch.full((D,), ls_init)
    state[f"{prefix}.ls2.gamma"] = torch.full((D,), ls_init)
    if qk_norm:
        for n in ("q_norm", "k_norm"):
            state[f"{prefix}.attn.{n}.weight"] = torch.ones(cfg.head_dim)
            state[f"{prefix}.attn.{n}.bias"] = torch.zeros(cfg.head_dim)


def init_aggregator_state(cfg: AggregatorConfig, seed: int = 0, sensitised: bool = False) -> State:
    """Random-init aggregator parameters under the upstream key names (Appendix A.6, prefix stripped).

    `sensitised=True` is the parity-stress weight set of SURVEY.md §7: LayerScale ~ U(0.1, 1),
    non-trivial camera/register tokens, QK-norm affine != (1, 0), non-zero biases and LayerNorm affine —
    with the default init the 48 aggregator blocks perturb the stream by ~1 % each and bugs hide.
    """
    gen = torch.Generator().manual_seed(seed)
    D = cfg.embed_dim
    g = cfg.img_size // cfg.patch_size
    s: State = {}
    s["camera_token"] = torch.randn(1, 2, 1, D, generator=gen) * 1e-6
    s["register_token"] = torch.randn(1, 2, cfg.num_register_tokens, D, generator=gen) * 1e-6
    # DINOv2 stem
    s["patch_embed.cls_token"] = torch.randn(1, 1, D, generator=gen) * 1e-6
    s["patch_embed.pos_embed"] = _trunc_normal(gen, (1, 1 + g * g, D), 0.02)
    s["patch_embed.register_tokens"] = torch.randn(1, cfg.num_register_tokens, D, generator=gen) * 1e-6
    s["patch_embed.mask_token"] = torch.zeros(1, D)
    fan_in = 3 * cfg.patch_size * cfg.patch_size
    bound = 1.0 / math.sqrt(fan_in)
    s["patch_embed.patch_embed.proj.weight"] = (torch.rand(D, 3, cfg.patch_size, cfg.patch_size, generator=gen) * 2 - 1) * bound
    s["patch_embed.patch_embed.proj.bias"] = (torch.rand(D, generator=gen) * 2 - 1) * bound
    for i in range(cfg.stem_depth):
        _block_params(s, f"patch_embed.blocks.{i}", cfg, gen, vit_init=True, qk_norm=False, ls_init=cfg.stem_init_values)
    s["patch_embed.norm.weight"], s["patch_embed.norm.bias"] = torch.ones(D), torch.zeros(D)
    for kind in ("frame_blocks", "global_blocks"):
        for i in range(cfg.depth):
            _block_params(s, f"{kind}.{i}", cfg, gen, vit_init=False, qk_norm=True, ls_init=cfg.init_values)
    if sensitised:
        for k in list(s.keys()):
            t = s[k]
            if k.endswith("gamma") and not k.startswith("patch_embed"):
                s[k] = torch.rand(t.shape, generator=gen) * 0.9 + 0.1
            elif k.endswith("gamma"):
                s[k] = torch.rand(t.shape, generator=gen) * 0.5 + 0.5
            elif k in ("camera_token", "register_token", "patch_embed.cls_token", "patch_embed.register_tokens"):
                s[k] = torch.randn(t.shape, generator=gen) * 0.5
            elif ("norm" in k) and k.endswith("weight"):
                s[k] = 1.0 + 0.3 * torch.randn(t.shape, generator=gen)
            elif ("norm" in k) and k.endswith("bias"):
                s[k] = 0.2 * torch.randn(t.shape, generator=gen)
            elif k.endswith("bias") and t.abs().max() == 0:
                s[k] = 0.05 * torch.randn(t.shape, generator=gen)
            elif k.endswith("weight") and k.startswith("patch_embed.blocks"):
                # trunc_normal(0.02) stem weights make the stem nearly linear; scale to nn.Linear-like size
                s[k] = t * (1.0 / math.sqrt(t.shape[1]) / 0.02) * 0.8
    return s


# --------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------

def rope_tables(max_pos: int, head_dim: int, freq: float = 100.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin tables [max_pos+1, head_dim//4] (Appendix A.4): each spatial axis owns head_dim/2 features,
    rotated in pairs (j, j + head_dim/4) with inv_freq_j = freq^(-2j / (head_dim/2))."""
    half = head_dim // 2
    j = torch.arange(0, half, 2, dtype=torch.float32)
    inv_freq = 1.0 / (freq ** (j / half))
    ang = torch.arange(max_pos + 1, dtype=torch.float32)[:, None] * inv_freq[None, :]
    return ang.cos(), ang.sin()


def _rope_1d(t: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """t [..., N, half]; pos [..., N] integer positions. out = t*cos + rotate_half(t)*sin, cos = cat(theta, theta)."""
    c = torch.cat([cos[pos], cos[pos]], dim=-1)
    s = torch.cat([sin[pos], sin[pos]], dim=-1)
    h = t.shape[-1] // 2
    rot = torch.cat([-t[..., h:], t[..., :h]], dim=-1)
    return t * c + rot * s


def rope_2d(t: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """t [B, heads, N, head_dim]; pos [B, N, 2] = (y, x). First half of the head rotates with y, second with x."""
    half = t.shape[-1] // 2
    py = pos[..., 0][:, None, :]
    px = pos[..., 1][:, None, :]
    return torch.cat([_rope_1d(t[..., :half], py, cos, sin), _rope_1d(t[..., half:], px, cos, sin)], dim=-1)


def _attention(s: State, p: str, x: torch.Tensor, heads: int, pos, tables) -> torch.Tensor:
    B, N, C = x.shape
    qkv = F.linear(x, s[f"{p}.qkv.weight"], s[f"{p}.qkv.bias"]).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if f"{p}.q_norm.weight" in s:
        hd = C // heads
        q = F.layer_norm(q, (hd,), s[f"{p}.q_norm.weight"], s[f"{p}.q_norm.bias"], 1e-5)
        k = F.layer_norm(k, (hd,), s[f"{p}.k_norm.weight"], s[f"{p}.k_norm.bias"], 1e-5)
    if pos is not None:
        q = rope_2d(q, pos, *tables)
        k = rope_2d(k, pos, *tables)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, s[f"{p}.proj.weight"], s[f"{p}.proj.bias"])


def block_forward(s: State, p: str, x: torch.Tensor, heads: int, eps: float, pos=None, tables=None) -> torch.Tensor:
    """Pre-norm transformer block with LayerScale: x += g1*attn(LN(x)); x += g2*mlp(LN(x))."""
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), s[f"{p}.norm1.weight"], s[f"{p}.norm1.bias"], eps)
    x = x + s[f"{p}.ls1.gamma"] * _attention(s, f"{p}.attn", h, heads, pos, tables)
    h = F.layer_norm(x, (C,), s[f"{p}.norm2.weight"], s[f"{p}.norm2.bias"], eps)
    h = F.linear(h, s[f"{p}.mlp.fc1.weight"], s[f"{p}.mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, s[f"{p}.mlp.fc2.weight"], s[f"{p}.mlp.fc2.bias"])
    return x + s[f"{p}.ls2.gamma"] * h


def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """DINOv2 `interpolate_pos_encoding` (interpolate_antialias=True, interpolate_offset=0.0): identity when the
    patch grid matches, else bicubic+antialias resize of the square grid in fp32."""
    n = pos_embed.shape[1] - 1
    m = int(round(math.sqrt(n)))
    if gh * gw == n and gh == gw:
        return pos_embed
    dim = pos_embed.shape[-1]
    patch = pos_embed[:, 1:].float().reshape(1, m, m, dim).permute(0, 3, 1, 2)
    patch = F.interpolate(patch, size=(gh, gw), mode="bicubic", antialias=True)
    patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, dim)
    return torch.cat([pos_embed[:, :1].float(), patch], dim=1)


def dinov2_forward(s: State, x: torch.Tensor, cfg: AggregatorConfig, prefix: str = "patch_embed") -> torch.Tensor:
    """DINOv2 ViT with registers on normalised images x [N,3,H,W] -> all normed tokens [N, 1+R+gh*gw, D]."""
    N, _, H, W = x.shape
    gh, gw = H // cfg.patch_size, W // cfg.patch_size
    t = F.conv2d(x, s[f"{prefix}.patch_embed.proj.weight"], s[f"{prefix}.patch_embed.proj.bias"], stride=cfg.patch_size)
    t = t.flatten(2).transpose(1, 2)                                   # [N, gh*gw, D] row-major patches
    t = torch.cat([s[f"{prefix}.cls_token"].expand(N, -1, -1), t], dim=1)
    t = t + interpolate_pos_embed(s[f"{prefix}.pos_embed"], gh, gw)
    t = torch.cat([t[:, :1], s[f"{prefix}.register_tokens"].expand(N, -1, -1), t[:, 1:]], dim=1)
    for i in range(cfg.stem_depth):
        t = block_forward(s, f"{prefix}.blocks.{i}", t, cfg.num_heads, 1e-6)
    return F.layer_norm(t, (cfg.embed_dim,), s[f"{prefix}.norm.weight"], s[f"{prefix}.norm.bias"], 1e-6)


def _special_tokens(tok: torch.Tensor, B: int, S: int) -> torch.Tensor:
    """`slice_expand_and_flatten`: index 0 for the first frame of a scene, index 1 for the others."""
    first = tok[:, 0:1].expand(B, 1, *tok.shape[2:])
    rest = tok[:, 1:2].expand(B, S - 1, *tok.shape[2:])
    return torch.cat([first, rest], dim=1).reshape(B * S, *tok.shape[2:])


def token_positions(gh: int, gw: int, n_special: int) -> torch.Tensor:
    """[n_special + gh*gw, 2] (y, x): patch (r, c) -> (r+1, c+1); special tokens -> (0, 0)."""
    ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    pos = torch.stack([ys.reshape(-1), xs.reshape(-1)], dim=-1) + 1
    return torch.cat([torch.zeros(n_special, 2, dtype=pos.dtype), pos], dim=0)


@torch.no_grad()
def aggregator_forward(s: State, images: torch.Tensor, cfg: AggregatorConfig, *, return_all: bool = True,
                       ) -> Tuple[List[torch.Tensor], int]:
    """images [B,S,3,H,W] in [0,1] -> (list of `depth` tensors [B,S,P,2*D], patch_start_idx).

    With return_all=False only the last list entry is materialised (the only one the reference consumes,
    src/models/vggt_qwen3_vlm.py:148)."""
    B, S, C_in, H, W = images.shape
    D = cfg.embed_dim
    images = images.float()
    dev = images.device   # the restatement is device-agnostic fp32 PyTorch: tests run the big cases with it on `cuda`
    mean = torch.tensor(_RESNET_MEAN, device=dev).view(1, 1, 3, 1, 1)
    std = torch.tensor(_RESNET_STD, device=dev).view(1, 1, 3, 1, 1)
    x = ((images - mean) / std).reshape(B * S, C_in, H, W)
    gh, gw = H // cfg.patch_size, W // cfg.patch_size
    n_special = cfg.patch_start_idx

    patch_tokens = dinov2_forward(s, x, cfg)[:, n_special:]                        # x_norm_patchtokens
    cam = _special_tokens(s["camera_token"], B, S)
    reg = _special_tokens(s["register_token"], B, S)
    tokens = torch.cat([cam, reg, patch_tokens], dim=1)                            # [B*S, P, D]
    P = tokens.shape[1]
    pos = token_positions(gh, gw, n_special).to(dev)[None].expand(B * S, -1, -1)
    tables = tuple(t.to(dev) for t in rope_tables(max(gh, gw), cfg.head_dim, cfg.rope_freq))

    outs: List[torch.Tensor] = []
    for i in range(cfg.depth):
        inter: Dict[str, torch.Tensor] = {}
        for kind in cfg.aa_order:
            if kind == "frame":
                tokens = block_forward(s, f"frame_blocks.{i}", tokens.reshape(B * S, P, D), cfg.num_heads, 1e-5,
                                       pos.reshape(B * S, P, 2), tables)
                inter["frame"] = tokens.reshape(B, S, P, D)
            else:
                tokens = block_forward(s, f"global_blocks.{i}", tokens.reshape(B, S * P, D), cfg.num_heads, 1e-5,
                                       pos.reshape(B, S * P, 2), tables)
                inter["global"] = tokens.reshape(B, S, P, D)
        if return_all or i == cfg.depth - 1:
            outs.append(torch.cat([inter["frame"], inter["global"]], dim=-1))
    return outs, n_special


def aggregator_flops(cfg: AggregatorConfig, S: int, H: int = 518, W: int = 518) -> float:
    """Algorithmic FLOPs per frame: 2MNK per GEMM, 4*Nq*Nkv*D per attention, no pruning credit."""
    gh, gw = H // cfg.patch_size, W // cfg.patch_size
    P = cfg.patch_start_idx + gh * gw
    D, Hd = cfg.embed_dim, cfg.hidden_dim
    lin = 2.0 * P * (D * 3 * D + D * D + 2 * D * Hd)
    att = 4.0 * P * P * D
    conv = 2.0 * gh * gw * (3 * cfg.patch_size ** 2) * D
    return conv + (cfg.stem_depth + 2 * cfg.depth) * lin + (cfg.stem_depth + cfg.depth) * att + cfg.depth * att * S
