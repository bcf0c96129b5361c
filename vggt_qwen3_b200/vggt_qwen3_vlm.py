"""VGGT + Perceiver projector + Qwen3 wrapper with the reference's module API, on the sm_100a path.

Mirrors src/models/vggt_qwen3_vlm.py of the reference: `VisionLanguageConfig` (:15-23) and `VGGTQwen3VLM`
(:26-201) keep their attribute names (.tokenizer .text_model .vision_model .projector .geom_head .num_vis_tokens
.geom_tokens), method signatures (encode_images / encode_geom / forward) and checkpoint key prefixes
(projector.* vision_model.aggregator.* geom_head.* text_model.*).  Qwen3 itself stays on stock PyTorch /
transformers (out of scope); what changes is what runs underneath:

  encode_images : one vq_aggregator_forward call (frozen VGGT aggregator, only the rows the slice at :148-156
                  keeps are materialised) + vq_perceiver_forward;
  forward       : the python splice loop with its `.nonzero()` host sync (:190-195) becomes one
                  vq_splice_overwrite launch.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .aggregator import VGGT, Aggregator
from .projector_perceiver import PerceiverConfig, PerceiverProjector


@dataclass
class VisionLanguageConfig:
    text_model_name: str
    vision_ckpt_dir: str
    num_vis_tokens: int = 64
    geom_tokens: int = 0
    projector_cfg: Optional[PerceiverConfig] = None
    freeze_vision: bool = True
    dtype: str = "bfloat16"


class _MockVGGT(nn.Module):
    """`vision_ckpt_dir == "mock"`: zero tokens, no external weights (reference :113-126).  The reference's mock
    returns a bare tensor that its own `encode_images` cannot unpack; this one returns the `(list, idx)` tuple the
    call site expects, keeping the documented intent (a CPU-free smoke path)."""

    def __init__(self, tokens: int, embed_dim: int = 256) -> None:
        super().__init__()
        self.tokens = max(tokens, 1)
        self.embed_dim = embed_dim

    def aggregator(self, images: torch.Tensor):
        B = images.shape[0]
        return [torch.zeros(B, self.tokens, self.embed_dim, device=images.device, dtype=images.dtype)], 5


class VGGTQwen3VLM(nn.Module):
    def __init__(self, config: VisionLanguageConfig, text_model: Optional[nn.Module] = None, tokenizer=None,
                 vision_model: Optional[nn.Module] = None) -> None:
        """`text_model` / `tokenizer` / `vision_model` may be injected (tests, offline boxes); otherwise they are
        loaded exactly as the reference does (:31-42, :60-111)."""
        super().__init__()
        dtype = getattr(torch, config.dtype, torch.bfloat16) if isinstance(config.dtype, str) else torch.bfloat16
        if tokenizer is None or text_model is None:
            from transformers import AutoModelForCausalLM, AutoTokenizer

            tokenizer = tokenizer or AutoTokenizer.from_pretrained(config.text_model_name)
            added = 0
            if "<image>" not in tokenizer.get_vocab():
                added = tokenizer.add_tokens(["<image>"])
            if text_model is None:
                text_model = AutoModelForCausalLM.from_pretrained(config.text_model_name, torch_dtype=dtype,
                                                                  use_safetensors=True)
                if added:
                    text_model.resize_token_embeddings(len(tokenizer))
        self.tokenizer = tokenizer
        self.text_model = text_model
        self.vision_model = vision_model if vision_model is not None else self._load_vggt(
            config.vision_ckpt_dir, config.num_vis_tokens)
        for p in self.vision_model.parameters():
            p.requires_grad_(not config.freeze_vision)
        hidden = self.text_model.config.hidden_size
        self.projector = PerceiverProjector(config.projector_cfg or PerceiverConfig(),
                                            in_dim=self.vision_model.embed_dim, out_dim=hidden)
        self.geom_head = nn.Sequential(nn.Linear(37, hidden), nn.SiLU(), nn.Linear(hidden, hidden))  # R9+t3+K9+hist16
        self.num_vis_tokens = config.num_vis_tokens
        self.geom_tokens = config.geom_tokens

    def _load_vggt(self, ckpt_dir: str, num_vis_tokens: int) -> nn.Module:
        if ckpt_dir == "mock":
            return _MockVGGT(tokens=num_vis_tokens)
        model = VGGT(img_size=518, patch_size=14, embed_dim=1024, enable_camera=True, enable_point=True,
                     enable_depth=True, enable_track=True)
        ckpt = Path(ckpt_dir) / "vggt_1B_commercial.pt"
        if ckpt.exists():
            state = torch.load(ckpt, map_location="cpu")
            if isinstance(state, dict):
                state = state.get("model", state.get("state_dict", state))
            model.load_state_dict(state, strict=False)  # prediction-head keys are not part of this path
        else:
            print(f"warning: {ckpt} not found, VGGT aggregator keeps its random initialisation")
        model = model.to(dtype=torch.bfloat16)
        model.eval()
        model.embed_dim = 2048
        return model

    # ------------------------------------------------------------------ hot path
    @staticmethod
    def prepare_views(views_u8: torch.Tensor, image_size: int) -> torch.Tensor:
        """The reference's per-view transform (src/dataio/collate_multiview.py:12-19: Resize(image_size, BICUBIC) +
        CenterCrop + ToTensor) on the GPU, for uint8 views [B, V, H, W, 3] as decoded: views that already have the
        target size are passed through as uint8 (the patch kernel applies x / 255 itself, nothing is materialised);
        others are resized / cropped bit-exactly (ops.preprocess_views) to float [B, V, 3, image_size, image_size]."""
        if views_u8.dtype != torch.uint8 or views_u8.dim() != 5 or views_u8.shape[-1] != 3:
            raise ValueError("prepare_views expects uint8 views [B, V, H, W, 3]")
        if views_u8.shape[2] == image_size and views_u8.shape[3] == image_size:
            return views_u8
        return ops.preprocess_views(views_u8, image_size)

    def _vision_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """[B, num_vis_tokens, embed_dim]: first tokens of the view-flattened last aggregator iteration (:144-156).
        `images`: float [B,V,3,H,W] in [0,1] like the reference, or uint8 views [B,V,H,W,3] (see prepare_views)."""
        agg_mod = self.vision_model.aggregator
        if isinstance(agg_mod, Aggregator):
            return agg_mod.encode_tokens(images, self.num_vis_tokens)
        if images.dtype == torch.uint8:
            images = images.permute(0, 1, 4, 2, 3).float() / 255.0
        tokens_list, _ = agg_mod(images.to(dtype=torch.bfloat16))
        agg = tokens_list[-1]
        if agg.dim() == 4:
            agg = agg.reshape(images.shape[0], -1, agg.shape[-1])
        return agg[:, : self.num_vis_tokens, :]

    def _projector_input(self, tokens: torch.Tensor) -> torch.Tensor:
        """The reference casts the aggregator tokens to the projector's parameter dtype (:159-160).  The sm_100a projector
        consumes bf16 tokens as they come out of the aggregator kernels (a cast to f32 and back would be two ATen
        launches for nothing); any other projector module gets the reference's cast."""
        if isinstance(self.projector, PerceiverProjector):
            return tokens
        return tokens.to(dtype=self.projector.in_proj.weight.dtype)

    @torch.no_grad()
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,V,C,H,W] -> [B, num_vis_tokens, hidden]; like the reference the whole call is under no_grad."""
        return self.projector(self._projector_input(self._vision_tokens(images)))

    def encode_images_trainable(self, images: torch.Tensor) -> torch.Tensor:
        """Same tokens, but the projector runs under autograd (the frozen aggregator stays under no_grad), so that
        stage-1 training actually reaches the Perceiver; the reference's `encode_images` cannot."""
        return self.project_tokens(self.encode_image_tokens(images))

    @torch.no_grad()
    def encode_image_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """The frozen half of `encode_images_trainable` (reference :140-160): aggregator tokens [B, num_vis_tokens,
        2048] as the projector consumes them (bf16 for the sm_100a projector).  Independent of the projector weights, so a data-parallel trainer can
        run it for step i+1 while the gradient all-reduce of step i is still in flight (bench.py does)."""
        return self._projector_input(self._vision_tokens(images))

    def project_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        """The trainable half (reference :162): `self.projector(tokens)` under autograd."""
        return self.projector(tokens)

    _GEOM_KEYS = (("R", "t"), ("t", "R"), ("K", "R"), ("depth_hist", "R"))   # (key, fallback shape donor), reference :168-173

    def _geom_vector(self, geom_token: Optional[Dict[str, torch.Tensor]]) -> Optional[torch.Tensor]:
        """geom_head(mean over views of cat(R, t, K, depth_hist)) -> f32 [B, hidden] (reference :164-176), on
        libvq_b200.so (vq_geom_head_forward, differentiable w.r.t. the geom_head parameters); None when bypassed."""
        if not geom_token or self.geom_tokens == 0:
            return None
        parts, widths = [], []
        for key, like in self._GEOM_KEYS:
            if key in geom_token:
                parts.append(geom_token[key])
                widths.append(geom_token[key].shape[-1])
            else:   # the reference substitutes zeros shaped like another entry
                parts.append(None)
                widths.append(geom_token[like].shape[-1])
        donor = next(p for p in parts if p is not None)
        if donor.dim() != 3:
            raise ValueError("geom_token entries must be [B, V, features]")
        lin1, lin2 = self.geom_head[0], self.geom_head[2]
        if sum(widths) != lin1.in_features:
            raise RuntimeError(f"geom features have width {sum(widths)}, geom_head expects {lin1.in_features}")
        return _GeomHeadFunction.apply(parts, widths, donor.shape[0], donor.shape[1], lin1.weight, lin1.bias,
                                       lin2.weight, lin2.bias)

    def encode_geom(self, geom_token: Optional[Dict[str, torch.Tensor]]) -> Optional[torch.Tensor]:
        """[B, geom_tokens, hidden] from per-view (R 9 | t 3 | K 9 | depth_hist 16) features averaged over views
        (reference :164-177); None when there is no geometry input or geom_tokens == 0."""
        y = self._geom_vector(geom_token)
        if y is None:
            return None
        return y.to(self.geom_head[0].weight.dtype).unsqueeze(1).expand(-1, self.geom_tokens, -1)

    def build_inputs_embeds(self, images: torch.Tensor, geom_token, input_ids: torch.Tensor,
                            trainable_projector: bool = False) -> torch.Tensor:
        """encode + splice: inputs_embeds with [geom, vis] written over every <image> position (:187-195).  The span
        `cat([geom.expand(geom_tokens), vis], 1)` is never materialised: one vq_splice_features launch reads both."""
        vis = self.encode_images_trainable(images) if trainable_projector else self.encode_images(images)
        geom = self._geom_vector(geom_token)
        inputs_embeds = self.text_model.get_input_embeddings()(input_ids)
        image_id = self.tokenizer.convert_tokens_to_ids("<image>")
        n_geom = self.geom_tokens if geom is not None else 0
        needs_grad = torch.is_grad_enabled() and (vis.requires_grad or inputs_embeds.requires_grad or
                                                  (geom is not None and geom.requires_grad))
        if needs_grad:
            return _SpliceFunction.apply(inputs_embeds, input_ids, geom, vis, image_id, n_geom)
        return ops.splice_features(inputs_embeds.contiguous(), input_ids, geom, n_geom, vis, image_id)

    def forward(self, images, geom_token, input_ids, attention_mask, labels) -> torch.Tensor:
        inputs_embeds = self.build_inputs_embeds(images, geom_token, input_ids)
        outputs = self.text_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels)
        return outputs.loss


class _GeomHeadFunction(torch.autograd.Function):
    """geom_head on the C ABI: forward vq_geom_head_forward, backward vq_geom_head_backward (parameter gradients only:
    the geometry features are data)."""

    @staticmethod
    def forward(ctx, parts, widths, B, V, w1, b1, w2, b2):
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        y, pooled, pre, act = ops.geom_head_forward([None if p is None else p.detach() for p in parts], widths,
                                                    f32(w1), f32(b1), f32(w2), f32(b2), B, V)
        ctx.save_for_backward(pooled, pre, act, w2)
        ctx.dtypes = (w1.dtype, b1.dtype, w2.dtype, b2.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        pooled, pre, act, w2 = ctx.saved_tensors
        grads = ops.geom_head_backward(dy, pooled, pre, act, w2.detach().to(torch.float32).contiguous())
        return (None, None, None, None, *[g.to(dt) for g, dt in zip(grads, ctx.dtypes)])


class _SpliceFunction(torch.autograd.Function):
    """Differentiable fused splice: forward vq_splice_features on a copy of inputs_embeds, backward
    vq_splice_features_backward (one launch: written rows route their gradient to the feature they came from and are
    zeroed in d(inputs_embeds)); no host synchronisation in either direction."""

    @staticmethod
    def forward(ctx, inputs_embeds, input_ids, geom, vis, image_id, n_geom):
        out = inputs_embeds.detach().clone().contiguous()
        ops.splice_features(out, input_ids, None if geom is None else geom.detach(), n_geom, vis.detach(), image_id)
        ctx.save_for_backward(input_ids)
        ctx.image_id, ctx.n_geom, ctx.n_vis = image_id, n_geom, vis.shape[1]
        ctx.vis_dtype = vis.dtype
        ctx.has_geom = geom is not None
        return out

    @staticmethod
    def backward(ctx, grad):
        (input_ids,) = ctx.saved_tensors
        need_emb, _, need_geom, need_vis = ctx.needs_input_grad[:4]
        d_emb, d_geom, d_vis = ops.splice_features_backward(
            grad, input_ids, ctx.n_geom, ctx.n_vis, ctx.image_id, need_embeds=need_emb,
            need_geom=need_geom and ctx.has_geom, need_vis=need_vis)
        return (d_emb, None, d_geom, None if d_vis is None else d_vis.to(ctx.vis_dtype), None, None)
