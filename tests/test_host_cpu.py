"""CPU tests of the host-side mirror of the reference interface: config dataclasses, state-dict key layout of the
module mirrors, weight-packing helpers (pos-embed interpolation, RoPE tables) against the oracle, and the FLOP model
bench.py reports with."""
import dataclasses
import sys
from pathlib import Path

import pytest
import torch

from oracle import aggregator_ref as A
from oracle import perceiver_ref as P

REF = Path("/root/reference")


def _import_reference(modname: str):
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REF))
    try:
        return __import__(modname, fromlist=["_"])
    finally:
        sys.path.remove(str(REF))


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not mounted")
def test_config_dataclasses_mirror_the_reference_field_for_field():
    ref_vlm = _import_reference("src.models.vggt_qwen3_vlm")
    ref_proj = _import_reference("src.models.projector_perceiver")
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig
    from vggt_qwen3_b200.vggt_qwen3_vlm import VisionLanguageConfig

    def fields(cls):
        out = []
        for f in dataclasses.fields(cls):
            default = None if f.default is dataclasses.MISSING else f.default
            out.append((f.name, default if not dataclasses.is_dataclass(default) else "dataclass"))
        return out

    assert fields(PerceiverConfig) == fields(ref_proj.PerceiverConfig)
    assert fields(VisionLanguageConfig) == fields(ref_vlm.VisionLanguageConfig)


def test_aggregator_mirror_has_the_upstream_state_dict_layout():
    """the keys a `vggt_1B_commercial.pt` checkpoint carries under `aggregator.`."""
    from vggt_qwen3_b200.aggregator import Aggregator

    agg = Aggregator(img_size=84, embed_dim=128, num_heads=2, depth=2, stem_depth=2)
    keys = set(agg.state_dict().keys())
    oracle_keys = set(A.init_aggregator_state(A.AggregatorConfig(img_size=84, embed_dim=128, num_heads=2, depth=2,
                                                                 stem_depth=2), seed=0).keys())
    assert keys == oracle_keys
    for k in ("camera_token", "register_token", "patch_embed.cls_token", "patch_embed.pos_embed",
              "patch_embed.register_tokens", "patch_embed.patch_embed.proj.weight", "patch_embed.norm.weight",
              "patch_embed.blocks.0.attn.qkv.weight", "patch_embed.blocks.1.ls2.gamma", "patch_embed.blocks.0.mlp.fc1.bias",
              "frame_blocks.0.attn.q_norm.weight", "global_blocks.1.attn.k_norm.bias", "global_blocks.0.mlp.fc2.weight"):
        assert k in keys, k
    sd = agg.state_dict()
    assert tuple(sd["camera_token"].shape) == (1, 2, 1, 128) and tuple(sd["register_token"].shape) == (1, 2, 4, 128)
    assert tuple(sd["patch_embed.patch_embed.proj.weight"].shape) == (128, 3, 14, 14)
    assert tuple(sd["patch_embed.pos_embed"].shape) == (1, 1 + 36, 128)
    assert tuple(sd["frame_blocks.0.attn.qkv.weight"].shape) == (384, 128)
    assert agg.patch_start_idx == 5 and agg.Kp == 592   # 3*14*14 = 588 padded to a multiple of 8 for 16-byte rows


def test_weight_packing_helpers_match_the_oracle():
    from vggt_qwen3_b200.aggregator import Aggregator, _rope_tables

    cos, sin = _rope_tables(37, 64, 100.0)
    rc, rs = A.rope_tables(37, 64, 100.0)
    assert cos.shape == (38, 16) and torch.allclose(cos, rc, atol=2e-6) and torch.allclose(sin, rs, atol=2e-6)
    assert torch.equal(cos[0], torch.ones(16)) and torch.equal(sin[0], torch.zeros(16))   # specials: identity rotation
    # stage1_3d.yaml trains at image_size 448 (32 x 32 patches): the 37 x 37 DINOv2 grid is interpolated
    agg = Aggregator(img_size=518, embed_dim=128, num_heads=2, depth=1, stem_depth=1)
    ours = agg._interpolated_pos_embed(32, 32)
    ref = A.interpolate_pos_embed(agg.patch_embed.pos_embed.detach().float(), 32, 32)
    assert ours.shape[-2] == 1 + 32 * 32
    assert torch.allclose(ours.reshape(ref.shape).float().cpu(), ref, atol=1e-6)
    same = agg._interpolated_pos_embed(37, 37)
    assert torch.equal(same.reshape(1, 1370, 128).float().cpu(), agg.patch_embed.pos_embed.detach().float())


def test_flop_model_matches_the_survey():
    """3048.0 / 4347.1 / 8801.2 GF per frame at S = 1 / 8 / 32, projector 315.68 GF per sample."""
    cfg = A.AggregatorConfig()
    for S, gf in ((1, 3048.0), (8, 4347.1), (32, 8801.2)):
        assert abs(A.aggregator_flops(cfg, S) / 1e9 - gf) < 0.6, S
    assert abs(P.perceiver_flops(P.PerceiverShape(), 128) / 1e9 - 315.68) < 0.05
    # bench.py totals: c2 31.96 TF, c3 140.37 TF, c4 281.95 TF per step per GPU
    c2 = 8 * A.aggregator_flops(cfg, 1) + 3 * 8 * P.perceiver_flops(P.PerceiverShape(), 128)
    c3 = 32 * A.aggregator_flops(cfg, 8) + 4 * P.perceiver_flops(P.PerceiverShape(), 128)
    c4 = 32 * A.aggregator_flops(cfg, 32) + 1 * P.perceiver_flops(P.PerceiverShape(), 128)
    assert abs(c2 / 1e12 - 31.96) < 0.02 and abs(c3 / 1e12 - 140.37) < 0.05 and abs(c4 / 1e12 - 281.95) < 0.05


def test_mock_vision_path_needs_no_weights():
    """`vision_ckpt_dir == "mock"` (reference :113-126): the module builds without external files and exposes the
    attributes the reference's callers read."""
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig
    from vggt_qwen3_b200.vggt_qwen3_vlm import VGGTQwen3VLM, VisionLanguageConfig

    class _Text(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = type("C", (), {"hidden_size": 96})()
            self.emb = torch.nn.Embedding(10, 96)

        def get_input_embeddings(self):
            return self.emb

    class _Tok:
        def get_vocab(self):
            return {"<image>": 3}

        def convert_tokens_to_ids(self, t):
            return 3

    cfg = VisionLanguageConfig(text_model_name="unused", vision_ckpt_dir="mock", num_vis_tokens=16, geom_tokens=2,
                               projector_cfg=PerceiverConfig(64, 16, 2, 1, 128, 0.0))
    m = VGGTQwen3VLM(cfg, text_model=_Text(), tokenizer=_Tok())
    for attr in ("tokenizer", "text_model", "vision_model", "projector", "geom_head", "num_vis_tokens", "geom_tokens"):
        assert hasattr(m, attr), attr
    assert m.num_vis_tokens == 16 and m.geom_tokens == 2
    names = [n for n, _ in m.named_parameters()]
    assert any(n.startswith("projector.") for n in names) and any(n.startswith("geom_head.") for n in names)
    geom = {"R": torch.zeros(1, 2, 9), "t": torch.zeros(1, 2, 3), "K": torch.zeros(1, 2, 9), "depth_hist": torch.zeros(1, 2, 16)}
    with pytest.raises(RuntimeError, match="no CPU implementation"):   # geom_head runs on libvq_b200.so only
        m.encode_geom(geom)
    assert m.encode_geom(None) is None and m.encode_geom({}) is None
