"""CPU tests: the oracle restatements against the committed golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import math
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import aggregator_ref as A
from oracle import perceiver_ref as P

GOLD = Path(__file__).resolve().parent / "golden"
REF = Path("/root/reference")


def _close(a, b, rtol=2e-4, what=""):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-12
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def _load_perceiver_golden():
    z = np.load(GOLD / "perceiver_ref_small.npz")
    state = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return z, state, grads


def test_perceiver_forward_matches_reference_golden():
    z, state, _ = _load_perceiver_golden()
    heads = int(z["cfg"][2])
    out = P.perceiver_forward(state, torch.from_numpy(z["tokens"]), heads)
    _close(out, z["out"], 1e-5, "perceiver forward")


def test_perceiver_backward_matches_reference_golden():
    z, state, grads = _load_perceiver_golden()
    heads = int(z["cfg"][2])
    out, g, gt = P.perceiver_forward_backward(state, torch.from_numpy(z["tokens"]), torch.from_numpy(z["grad_out"]),
                                              heads, dtype=torch.float64)
    _close(out, z["out"], 1e-5, "forward")
    _close(gt, z["grad_tokens"], 2e-4, "grad tokens")
    assert set(g) == set(grads)
    for k in grads:
        _close(g[k], grads[k], 2e-4, f"grad {k}")


def test_perceiver_state_keys_and_param_count_full_size():
    shape = P.PerceiverShape()
    # parameter count of the reference class at perceiver_small / in 2048 / out 2560
    E, F, L = shape.latent_dim, shape.ffn_dim, shape.num_layers
    per_layer = 3 * E * E + 3 * E + E * E + E + 2 * E * F + F + E + 4 * E
    total = shape.num_latents * E + (shape.in_dim * E + E) + L * per_layer + (E * shape.out_dim + shape.out_dim)
    assert total == 1_227_684_352
    assert abs(P.perceiver_flops(shape, 128) - 315.68e9) < 0.01e9


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not mounted")
def test_perceiver_matches_live_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REF))
    try:
        from src.models.projector_perceiver import PerceiverConfig, PerceiverProjector
    finally:
        sys.path.remove(str(REF))
    torch.manual_seed(7)
    cfg = PerceiverConfig(latent_dim=192, num_latents=32, num_heads=3, num_layers=3, ffn_dim=384, dropout=0.0)
    m = PerceiverProjector(cfg, in_dim=48, out_dim=80).double()
    tokens = torch.randn(2, 40, 48, dtype=torch.float64, requires_grad=True)
    out = m(tokens)
    go = torch.randn_like(out)
    out.backward(go)
    state = {k: v.detach() for k, v in m.named_parameters()}
    o2, g2, gt2 = P.perceiver_forward_backward(state, tokens.detach(), go, cfg.num_heads, dtype=torch.float64)
    _close(o2, out.detach(), 1e-10, "forward")
    _close(gt2, tokens.grad, 1e-9, "grad tokens")
    for k, p in m.named_parameters():
        _close(g2[k], p.grad, 1e-9, f"grad {k}")
    # the restated initialiser produces the reference's key set and shapes
    init = P.init_perceiver_state(P.PerceiverShape(192, 32, 3, 3, 384, 0.0, 48, 80))
    assert {k: tuple(v.shape) for k, v in init.items()} == {k: tuple(v.shape) for k, v in state.items()}


def test_encode_images_token_order_golden():
    z = np.load(GOLD / "encode_images_ref.npz")
    B, V, Pn, C, n = (int(v) for v in z["shape"])
    b = torch.arange(B).view(B, 1, 1, 1) * 10000
    v = torch.arange(V).view(1, V, 1, 1) * 100
    p = torch.arange(Pn).view(1, 1, Pn, 1)
    last = (b + v + p).expand(B, V, Pn, C).to(torch.bfloat16)
    sel = P.select_vis_tokens(last, n)[..., :1].float()
    assert sel.shape == (B, n, 1)
    assert torch.equal(sel, torch.from_numpy(z["selected"]))  # exact: order and count
    # frame 0's tokens come first, then frame 1 continues
    assert sel[0, Pn - 1, 0] == Pn - 1 and sel[0, Pn, 0] == 100
    # rank-3 variant
    assert torch.equal(P.select_vis_tokens(last.reshape(B, V * Pn, C), n), P.select_vis_tokens(last, n))


def test_insert_splice_golden():
    z = np.load(GOLD / "insert_splice_ref.npz")
    emb, mask = P.splice_insert(torch.from_numpy(z["emb"]), torch.from_numpy(z["mask"]), torch.from_numpy(z["ids"]),
                                torch.from_numpy(z["vis"]), int(z["image_id"]))
    assert torch.equal(emb, torch.from_numpy(z["new_emb"]))
    assert torch.equal(mask, torch.from_numpy(z["new_mask"]))
    ids = torch.from_numpy(z["ids"]).clone()
    ids[ids == int(z["image_id"])] = 1
    e2, m2 = P.splice_insert(torch.from_numpy(z["emb"]), torch.from_numpy(z["mask"]), ids, torch.from_numpy(z["vis"]),
                             int(z["image_id"]))
    assert torch.equal(e2, torch.from_numpy(z["emb"])) and torch.equal(m2, torch.from_numpy(z["mask"]))


def _reference_overwrite_loop(inputs_embeds, input_ids, features, image_id):
    """The five reference lines (src/models/vggt_qwen3_vlm.py:191-195) on plain tensors, as the pin for the splice."""
    inputs_embeds = inputs_embeds.clone()
    image_positions = (input_ids == image_id).nonzero(as_tuple=False)
    for batch_idx, pos in image_positions:
        span = features[batch_idx]
        inputs_embeds[batch_idx, pos: pos + span.size(0), :] = span
    return inputs_embeds


def test_overwrite_splice_matches_reference_loop():
    g = torch.Generator().manual_seed(0)
    B, L, H, n = 3, 40, 8, 6
    ids = torch.randint(0, 50, (B, L), generator=g)
    ids[0, 3] = 99
    ids[1, 10] = 99
    ids[1, 13] = 99   # overlapping spans: the later one wins on the overlap
    ids[2, L - n] = 99  # exactly fits
    emb = torch.randn(B, L, H, generator=g)
    feats = torch.randn(B, n, H, generator=g)
    assert torch.equal(P.splice_overwrite(emb, ids, feats, 99), _reference_overwrite_loop(emb, ids, feats, 99))
    ids[2, L - 2] = 99  # does not fit: the reference raises
    with pytest.raises(RuntimeError):
        _reference_overwrite_loop(emb, ids, feats, 99)
    with pytest.raises(RuntimeError):
        P.splice_overwrite(emb, ids, feats, 99)


def _hf_to_upstream(z):
    """Map HF Dinov2WithRegisters weights onto the upstream `patch_embed.*` keys used by the oracle."""
    hf = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("hf/")}
    s = {
        "patch_embed.cls_token": hf["embeddings.cls_token"],
        "patch_embed.pos_embed": hf["embeddings.position_embeddings"],
        "patch_embed.register_tokens": hf["embeddings.register_tokens"],
        "patch_embed.mask_token": hf["embeddings.mask_token"],
        "patch_embed.patch_embed.proj.weight": hf["embeddings.patch_embeddings.projection.weight"],
        "patch_embed.patch_embed.proj.bias": hf["embeddings.patch_embeddings.projection.bias"],
        "patch_embed.norm.weight": hf["layernorm.weight"],
        "patch_embed.norm.bias": hf["layernorm.bias"],
    }
    n_layers = 1 + max(int(k.split(".")[2]) for k in hf if k.startswith("encoder.layer."))
    for i in range(n_layers):
        h, u = f"encoder.layer.{i}", f"patch_embed.blocks.{i}"
        a = f"{h}.attention.attention"
        s[f"{u}.attn.qkv.weight"] = torch.cat([hf[f"{a}.query.weight"], hf[f"{a}.key.weight"], hf[f"{a}.value.weight"]])
        s[f"{u}.attn.qkv.bias"] = torch.cat([hf[f"{a}.query.bias"], hf[f"{a}.key.bias"], hf[f"{a}.value.bias"]])
        s[f"{u}.attn.proj.weight"] = hf[f"{h}.attention.output.dense.weight"]
        s[f"{u}.attn.proj.bias"] = hf[f"{h}.attention.output.dense.bias"]
        for n in ("norm1", "norm2"):
            s[f"{u}.{n}.weight"], s[f"{u}.{n}.bias"] = hf[f"{h}.{n}.weight"], hf[f"{h}.{n}.bias"]
        s[f"{u}.ls1.gamma"] = hf[f"{h}.layer_scale1.lambda1"]
        s[f"{u}.ls2.gamma"] = hf[f"{h}.layer_scale2.lambda1"]
        for n in ("fc1", "fc2"):
            s[f"{u}.mlp.{n}.weight"], s[f"{u}.mlp.{n}.bias"] = hf[f"{h}.mlp.{n}.weight"], hf[f"{h}.mlp.{n}.bias"]
    return s, n_layers


def test_dinov2_stem_matches_hf_golden():
    z = np.load(GOLD / "dinov2_hf_tiny.npz")
    state, n_layers = _hf_to_upstream(z)
    cfg = A.AggregatorConfig(img_size=56, embed_dim=128, num_heads=2, stem_depth=n_layers, depth=1)
    y = A.dinov2_forward(state, torch.from_numpy(z["x"]), cfg)
    _close(y, z["y"], 2e-5, "DINOv2 stem vs HF Dinov2WithRegistersModel")


def test_rope_properties():
    cos, sin = A.rope_tables(37, 64)
    assert cos.shape == (38, 16)
    assert torch.allclose(cos[0], torch.ones(16)) and torch.allclose(sin[0], torch.zeros(16))  # specials: identity
    t = torch.randn(1, 2, 5, 64)
    pos = torch.tensor([[[0, 0], [1, 1], [3, 7], [37, 37], [2, 0]]])
    r = A.rope_2d(t, pos, cos, sin)
    assert torch.allclose(r[:, :, 0], t[:, :, 0])  # position (0,0) untouched
    assert torch.allclose(r.norm(dim=-1), t.norm(dim=-1), atol=1e-5)  # rotations preserve norm
    # relative-position property of rotary embeddings along one axis
    q, k = torch.randn(1, 1, 1, 64), torch.randn(1, 1, 1, 64)
    def dot(py, px, qy, qx):
        a = A.rope_2d(q, torch.tensor([[[py, px]]]), cos, sin)
        b = A.rope_2d(k, torch.tensor([[[qy, qx]]]), cos, sin)
        return (a * b).sum()
    assert torch.allclose(dot(5, 9, 2, 4), dot(15, 19, 12, 14), atol=1e-4)


def test_aggregator_oracle_shapes_and_flops():
    cfg = A.AggregatorConfig(img_size=56, embed_dim=128, num_heads=2, depth=2, stem_depth=1)
    st = A.init_aggregator_state(cfg, 0)
    outs, idx = A.aggregator_forward(st, torch.rand(2, 3, 3, 56, 56), cfg)
    assert idx == 5 and len(outs) == 2 and outs[-1].shape == (2, 3, 21, 256)
    last, _ = A.aggregator_forward(st, torch.rand(2, 3, 3, 56, 56, generator=torch.Generator().manual_seed(0)), cfg,
                                   return_all=False)
    assert len(last) == 1
    full = A.AggregatorConfig()
    assert abs(A.aggregator_flops(full, 1) - 3048.0e9) < 0.1e9
    assert abs(A.aggregator_flops(full, 8) - 4347.1e9) < 0.1e9
    assert abs(A.aggregator_flops(full, 32) - 8801.2e9) < 0.1e9
    # global attention mixes frames, frame attention does not: perturbing frame 1 changes frame 0 only via global
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(1, 2, 3, 56, 56, generator=g)
    o1, _ = A.aggregator_forward(st, imgs, cfg)
    imgs2 = imgs.clone()
    imgs2[:, 1] = torch.rand(3, 56, 56, generator=g)
    o2, _ = A.aggregator_forward(st, imgs2, cfg)
    D = cfg.embed_dim
    assert torch.equal(o1[0][:, 0, :, :D], o2[0][:, 0, :, :D])       # frame-block half of iteration 0, frame 0
    assert not torch.equal(o1[0][:, 0, :, D:], o2[0][:, 0, :, D:])   # global half sees frame 1
