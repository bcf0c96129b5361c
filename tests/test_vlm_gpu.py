"""GPU tests through the reference-facing module API (VGGTQwen3VLM): encode_images, the training splice in
forward(), and one full-size (518 px, 24+48 blocks, D=1024) frame against the oracle."""
import types

import pytest
import torch
import torch.nn as nn

from oracle import aggregator_ref as A
from oracle import perceiver_ref as P
from tests.parity import assert_token_parity

pytestmark = pytest.mark.gpu
IMAGE_ID = 7


class _Tok:
    def get_vocab(self):
        return {"<image>": IMAGE_ID}

    def convert_tokens_to_ids(self, token):
        assert token == "<image>"
        return IMAGE_ID


class _Text(nn.Module):
    """Stand-in for Qwen3: embedding table + a loss that depends on every inputs_embeds element."""

    def __init__(self, hidden: int, vocab: int = 64):
        super().__init__()
        self.config = types.SimpleNamespace(hidden_size=hidden)
        self.embed = nn.Embedding(vocab, hidden)
        self.last_inputs_embeds = None

    def get_input_embeddings(self):
        return self.embed

    def forward(self, inputs_embeds, attention_mask, labels):
        self.last_inputs_embeds = inputs_embeds
        return types.SimpleNamespace(loss=(inputs_embeds.float() * attention_mask[..., None]).pow(2).mean())


def _small_vlm(num_vis_tokens=64, geom_tokens=0, hidden=192):
    from vggt_qwen3_b200.aggregator import VGGT
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig
    from vggt_qwen3_b200.vggt_qwen3_vlm import VGGTQwen3VLM, VisionLanguageConfig

    cfg = A.AggregatorConfig(img_size=84, embed_dim=128, num_heads=2, depth=2, stem_depth=2)
    state = A.init_aggregator_state(cfg, seed=0, sensitised=True)
    vggt = VGGT(img_size=84, patch_size=14, embed_dim=128, num_heads=2, depth=2, stem_depth=2)
    vggt.aggregator.load_state_dict(state, strict=True)
    vcfg = VisionLanguageConfig("stub", "unused", num_vis_tokens=num_vis_tokens, geom_tokens=geom_tokens,
                                projector_cfg=PerceiverConfig(256, 64, 2, 2, 512, 0.0))
    model = VGGTQwen3VLM(vcfg, text_model=_Text(hidden), tokenizer=_Tok(), vision_model=vggt)
    shape = P.PerceiverShape(256, 64, 2, 2, 512, 0.0, 256, hidden)
    pstate = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
    pstate["latents"] = pstate["latents"] * 25.0
    model.projector.load_state_dict(pstate, strict=True)
    return model.cuda().eval(), cfg, state, shape, pstate


def test_encode_images_matches_oracle_and_contract():
    model, cfg, state, shape, pstate = _small_vlm()
    assert model.vision_model.embed_dim == 256 and model.projector.in_proj.weight.dtype == torch.float32
    images = torch.rand(2, 3, 3, 84, 84, generator=torch.Generator().manual_seed(1))
    ref_list, idx = A.aggregator_forward(state, images, cfg, return_all=False)
    ref = P.perceiver_forward(pstate, P.select_vis_tokens(ref_list[-1], 64), shape.num_heads)
    out = model.encode_images(images.cuda())
    assert out.shape == (2, 64, 192) and out.dtype == torch.float32 and not out.requires_grad
    assert_token_parity(out, ref, "encode_images")
    # the upstream aggregator call signature is preserved
    lst, start = model.vision_model.aggregator(images.cuda())
    assert start == 5 and lst[-1].shape == (2, 3, 41, 256)
    # trainable variant reaches the projector parameters, the frozen aggregator stays without grad
    out2 = model.encode_images_trainable(images.cuda())
    assert out2.requires_grad and torch.equal(out2.detach(), out)
    out2.sum().backward()
    assert all(p.grad is not None for p in model.projector.parameters())
    assert all(p.grad is None for p in model.vision_model.parameters())


def test_forward_splices_like_the_reference():
    model, cfg, state, shape, pstate = _small_vlm(geom_tokens=4)
    B, L = 2, 96
    g = torch.Generator().manual_seed(2)
    images = torch.rand(B, 2, 3, 84, 84, generator=g).cuda()
    ids = torch.randint(8, 64, (B, L), generator=g)
    ids[0, 3] = IMAGE_ID
    ids[1, 20] = IMAGE_ID
    ids = ids.cuda()
    mask = torch.ones(B, L, dtype=torch.long, device="cuda")
    geom = {k: torch.randn(B, 2, n, generator=g).cuda() for k, n in (("R", 9), ("t", 3), ("K", 9), ("depth_hist", 16))}
    loss = model(images, geom, ids, mask, labels=ids)
    assert torch.isfinite(loss)
    got = model.text_model.last_inputs_embeds
    vis = model.encode_images(images)
    geo = model.encode_geom(geom)
    feats = torch.cat([geo, vis], dim=1)
    assert feats.shape[1] == 4 + 64
    base = model.text_model.get_input_embeddings()(ids)
    want = P.splice_overwrite(base.detach().cpu(), ids.cpu(), feats.detach().cpu(), IMAGE_ID)
    assert torch.equal(got.detach().cpu(), want)
    # geom head restatement agrees with the module
    gstate = {"geom_head." + k: v.detach().cpu() for k, v in model.geom_head.state_dict().items()}
    gref = P.encode_geom(gstate, {k: v.cpu() for k, v in geom.items()}, 4)
    assert torch.allclose(geo.detach().cpu(), gref, atol=1e-5)
    # differentiable splice: gradients reach the projector through inputs_embeds
    emb = model.build_inputs_embeds(images, geom, ids, trainable_projector=True)
    emb.float().pow(2).sum().backward()
    assert model.projector.out_proj.weight.grad.abs().sum() > 0


def test_full_size_single_frame_matches_oracle():
    """c1 of BASELINE.json: one 518x518 image, full VGGT aggregator (24 DINOv2 + 24x2 blocks, D=1024) + perceiver_small
    -> 128 tokens, bf16 CUDA path vs the fp32 CPU oracle on the same seed-0 random-init weights."""
    from vggt_qwen3_b200.aggregator import Aggregator
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    cfg = A.AggregatorConfig()
    state = A.init_aggregator_state(cfg, seed=0)
    image = torch.rand(1, 1, 3, 518, 518, generator=torch.Generator().manual_seed(0))
    ref_list, _ = A.aggregator_forward(state, image, cfg, return_all=False)
    assert ref_list[-1].shape == (1, 1, 1374, 2048)
    agg = Aggregator()
    agg.load_state_dict(state, strict=True)
    agg = agg.cuda().eval()
    lst, _ = agg(image.cuda())
    assert lst[-1].shape == (1, 1, 1374, 2048)
    assert_token_parity(lst[-1], ref_list[-1], "full-size aggregator output")
    tokens = agg.encode_tokens(image.cuda(), 128)
    assert torch.equal(tokens, lst[-1].reshape(1, 1374, 2048)[:, :128])
    del agg, state
    shape = P.PerceiverShape()
    pstate = P.init_perceiver_state(shape, seed=0)
    ref = P.perceiver_forward(pstate, P.select_vis_tokens(ref_list[-1], 128), shape.num_heads)
    proj = PerceiverProjector(PerceiverConfig(), in_dim=2048, out_dim=2560)
    proj.load_state_dict(pstate, strict=True)
    proj = proj.cuda().eval()
    with torch.no_grad():
        out = proj(tokens)
    assert out.shape == (1, 128, 2560)
    assert_token_parity(out, ref, "full-size encode (128 tokens into Qwen3-4B embeds)")
