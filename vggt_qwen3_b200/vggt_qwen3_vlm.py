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
    def _vision_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """[B, num_vis_tokens, embed_dim]: first tokens of the view-flattened last aggregator iteration (:144-156)."""
        agg_mod = self.vision_model.aggregator
        if isinstance(agg_mod, Aggregator):
            return agg_mod.encode_tokens(images, self.num_vis_tokens)
        tokens_list, _ = agg_mod(images.to(dtype=torch.bfloat16))
        agg = tokens_list[-1]
        if agg.dim() == 4:
            agg = agg.reshape(images.shape[0], -1, agg.shape[-1])
        return agg[:, : self.num_vis_tokens, :]

    @torch.no_grad()
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,V,C,H,W] -> [B, num_vis_tokens, hidden]; like the reference the whole call is under no_grad."""
        agg = self._vision_tokens(images)
        return self.projector(agg.to(dtype=self.projector.in_proj.weight.dtype))

    def encode_images_trainable(self, images: torch.Tensor) -> torch.Tensor:
        """Same tokens, but the projector runs under autograd (the frozen aggregator stays under no_grad), so that
        stage-1 training actually reaches the Perceiver; the reference's `encode_images` cannot."""
        return self.project_tokens(self.encode_image_tokens(images))

    @torch.no_grad()
    def encode_image_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """The frozen half of `encode_images_trainable` (reference :140-160): aggregator tokens [B, num_vis_tokens,
        2048] in the projector's input dtype.  Independent of the projector weights, so a data-parallel trainer can
        run it for step i+1 while the gradient all-reduce of step i is still in flight (bench.py does)."""
        return self._vision_tokens(images).to(dtype=self.projector.in_proj.weight.dtype)

    def project_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        """The trainable half (reference :162): `self.projector(tokens)` under autograd."""
        return self.projector(tokens)

    def encode_geom(self, geom_token: Optional[Dict[str, torch.Tensor]]) -> Optional[torch.Tensor]:
        """[B, geom_tokens, hidden] from per-view (R 9 | t 3 | K 9 | depth_hist 16) features averaged over views
        (reference :164-177); None when there is no geometry input or geom_tokens == 0."""
        if not geom_token or self.geom_tokens == 0:
            return None
        parts = []
        for key, like in (("R", "t"), ("t", "R"), ("K", "R"), ("depth_hist", "R")):
            parts.append(geom_token[key] if key in geom_token else torch.zeros_like(geom_token[like]))
        pooled = torch.cat(parts, dim=-1).mean(dim=1)
        return self.geom_head(pooled).unsqueeze(1).expand(-1, self.geom_tokens, -1)

    def build_inputs_embeds(self, images: torch.Tensor, geom_token, input_ids: torch.Tensor,
                            trainable_projector: bool = False) -> torch.Tensor:
        """encode + splice: inputs_embeds with [geom, vis] written over every <image> position (:187-195)."""
        vis = self.encode_images_trainable(images) if trainable_projector else self.encode_images(images)
        geom = self.encode_geom(geom_token)
        features = vis if geom is None else torch.cat([geom.to(vis.dtype), vis], dim=1)
        inputs_embeds = self.text_model.get_input_embeddings()(input_ids)
        image_id = self.tokenizer.convert_tokens_to_ids("<image>")
        if features.requires_grad or inputs_embeds.requires_grad:
            return _SpliceFunction.apply(inputs_embeds, input_ids, features, image_id)
        return ops.splice_overwrite(inputs_embeds.contiguous(), input_ids, features, image_id)

    def forward(self, images, geom_token, input_ids, attention_mask, labels) -> torch.Tensor:
        inputs_embeds = self.build_inputs_embeds(images, geom_token, input_ids)
        outputs = self.text_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels)
        return outputs.loss


class _SpliceFunction(torch.autograd.Function):
    """Differentiable splice: d(features) gathers the rows written, d(inputs_embeds) is zero there."""

    @staticmethod
    def forward(ctx, inputs_embeds, input_ids, features, image_id):
        out = inputs_embeds.detach().clone().contiguous()
        ops.splice_overwrite(out, input_ids, features.detach(), image_id)
        ctx.save_for_backward(input_ids)
        ctx.image_id, ctx.n = image_id, features.shape[1]
        ctx.feat_dtype = features.dtype
        return out

    @staticmethod
    def backward(ctx, grad):
        (input_ids,) = ctx.saved_tensors
        B, L = input_ids.shape
        n = ctx.n
        # row r of sample b was written from features[b, r - pos] where pos is the last <image> position in (r-n, r]
        is_img = input_ids == ctx.image_id
        idx = torch.arange(L, device=grad.device).expand(B, L)
        last = torch.cummax(torch.where(is_img, idx, torch.full_like(idx, -1)), dim=1).values
        written = (last >= 0) & (idx - last < n)
        g_feat = torch.zeros(B, n, grad.shape[-1], device=grad.device, dtype=torch.float32)
        b_idx, r_idx = written.nonzero(as_tuple=True)
        g_feat.index_put_((b_idx, (idx - last)[b_idx, r_idx]), grad[b_idx, r_idx].float(), accumulate=True)
        g_emb = grad.masked_fill(written.unsqueeze(-1), 0)
        return g_emb, None, g_feat.to(ctx.feat_dtype), None
