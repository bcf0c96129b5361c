"""GPU parity of the one-tile Perceiver attention kernels (csrc/perceiver_attn.cu, through the C ABI) against an fp32
PyTorch statement of nn.MultiheadAttention's core and of its autograd backward
(src/models/projector_perceiver.py:33,45), with and without attention-probability dropout; and of the Perceiver
forward + backward run through the one-tile kernels against the multi-launch formulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, heads, Lq, T, head_dim)
SHAPES = [
    (2, 2, 64, 64, 64),      # the golden's geometry
    (3, 2, 64, 72, 64),      # T not a multiple of 64: keys 72..127 masked, rows 64..127 of the tile unused
    (2, 4, 128, 24, 128),    # short context
    (2, 2, 128, 128, 192),   # head_dim not a multiple of 128: last output piece is 64 wide
    (8, 8, 128, 128, 512),   # perceiver_small on the c2 step
]


def _inputs(B, H, Lq, T, hd, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    E = H * hd
    q = torch.randn(B * Lq, E, device="cuda", generator=g).mul_(1.5).bfloat16()
    kv = torch.randn(B * T, 2 * E, device="cuda", generator=g).bfloat16()
    datt = torch.randn(B * Lq, E, device="cuda", generator=g).bfloat16()
    return q, kv, datt


def _heads(x, B, L, H, hd, col0=0):
    """[B*L, ...] -> f32 [B, H, L, hd] taking columns col0 + h*hd."""
    return x[:, col0:col0 + H * hd].float().reshape(B, L, H, hd).permute(0, 2, 1, 3)


def _reference(q, kv, datt, probs_bf16, mask, p, B, H, Lq, T, hd):
    """fp32 forward from (q, k, v) and backward from the bf16 probabilities the kernel saved (what autograd would see)."""
    E = H * hd
    qh, kh, vh, doh = _heads(q, B, Lq, H, hd), _heads(kv, B, T, H, hd), _heads(kv, B, T, H, hd, E), _heads(datt, B, Lq, H, hd)
    scale = hd ** -0.5
    P = torch.softmax(scale * qh @ kh.transpose(-1, -2), dim=-1)
    Pb = probs_bf16[:, :T].float().reshape(B, H, Lq, T)
    keep = mask.float().reshape(B, H, Lq, T) / (1.0 - p) if mask is not None else torch.ones_like(Pb)
    Pd = (Pb * keep).bfloat16().float()
    att = Pd @ vh
    dPd = doh @ vh.transpose(-1, -2)
    dP = dPd * keep
    dS = scale * Pb * (dP - (dP * Pb).sum(-1, keepdim=True))
    dq, dk, dv = dS @ kh, dS.transpose(-1, -2) @ qh, Pd.transpose(-1, -2) @ doh
    merge = lambda x, L: x.permute(0, 2, 1, 3).reshape(B * L, E)
    return P, Pd, merge(att, Lq), merge(dq, Lq), torch.cat([merge(dk, T), merge(dv, T)], dim=1)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,H,Lq,T,hd", SHAPES)
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_one_tile_attention_matches_fp32(B, H, Lq, T, hd, p):
    from vggt_qwen3_b200 import ops

    E, Tp = H * hd, (T + 63) // 64 * 64
    assert ops.perceiver_attention_supported(Lq, T, E, H)
    q, kv, datt = _inputs(B, H, Lq, T, hd, seed=B + T + hd)
    drop = ops.dropout_desc(p, 1234, 8, 4 * 3 + 0) if p > 0 else None
    mask = ops.dropout_mask(B * H * Lq, T, Tp, p, 1234, 8, 4 * 3 + 0).bool() if p > 0 else None
    att, probs, dropped = ops.perceiver_attention_fwd(q, kv, B, H, drop)
    dq, dkv = ops.perceiver_attention_bwd(q, kv, probs, datt, B, H, drop)
    torch.cuda.synchronize()
    assert probs.shape == (B * H * Lq, Tp) and (dropped is None) == (p == 0)
    P, Pd, att_ref, dq_ref, dkv_ref = _reference(q, kv, datt, probs, mask, p, B, H, Lq, T, hd)
    # probabilities: bf16 rounding of an f32 softmax; padding columns are exact zeros
    assert (probs[:, :T].float().reshape(P.shape) - P).abs().max() <= 4e-3 + 8e-3 * P.max()
    assert (probs[:, T:] == 0).all()
    assert (probs[:, :T].float().sum(-1) - 1).abs().max() <= 2e-2
    if p > 0:
        # the dropped copy is exactly mask * bf16(P) / (1 - p) in bf16, zero where the site's mask says so
        assert torch.equal(dropped[:, :T].float().reshape(Pd.shape), Pd)
        assert (dropped[:, T:] == 0).all()
        kept = mask.float().mean().item()
        assert abs(kept - (1 - p)) < 0.02
    for name, got, ref in (("att", att, att_ref), ("dq", dq, dq_ref), ("dkv", dkv, dkv_ref)):
        assert torch.isfinite(got.float()).all(), name
        assert _rel(got, ref) <= 1e-2, f"{name}: rel L2 {_rel(got, ref):.3e}"


def test_rows_and_keys_outside_the_tile_are_untouched():
    """Lq = 64 and T = 24 leave half of the 128 x 128 tile unused: the kernels must neither write rows of the next
    sample nor let the next sample's keys (which the 128-row boxes do load) leak into the sums."""
    from vggt_qwen3_b200 import ops

    B, H, Lq, T, hd = 3, 2, 64, 24, 64
    q, kv, datt = _inputs(B, H, Lq, T, hd, seed=5)
    att, probs, _ = ops.perceiver_attention_fwd(q, kv, B, H)
    dq, dkv = ops.perceiver_attention_bwd(q, kv, probs, datt, B, H)
    # sample 1 computed alone must give the same bits as inside the batch
    s = slice(Lq, 2 * Lq)
    t = slice(T, 2 * T)
    att1, probs1, _ = ops.perceiver_attention_fwd(q[s].contiguous(), kv[t].contiguous(), 1, H)
    dq1, dkv1 = ops.perceiver_attention_bwd(q[s].contiguous(), kv[t].contiguous(), probs1, datt[s].contiguous(), 1, H)
    torch.cuda.synchronize()
    assert torch.equal(att[s], att1) and torch.equal(probs[H * Lq:2 * H * Lq], probs1)
    assert torch.equal(dq[s], dq1) and torch.equal(dkv[t], dkv1)


def test_unsupported_geometry_is_refused():
    from vggt_qwen3_b200 import ops

    assert not ops.perceiver_attention_supported(192, 128, 512, 4)   # more latents than one tile
    assert not ops.perceiver_attention_supported(128, 256, 512, 4)   # longer context than one tile
    assert not ops.perceiver_attention_supported(128, 128, 512, 16)  # head_dim 32
    q = torch.zeros(2 * 192, 512, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(2 * 128, 1024, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="one-tile attention"):
        ops.perceiver_attention_fwd(q, kv, 2, 4)


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_projector_same_through_both_formulations(dropout):
    """PerceiverProjector forward + backward with the one-tile kernels against the multi-launch formulation (batched
    GEMMs + row softmax): same masks, same saved probabilities, outputs and gradients equal to bf16 rounding noise."""
    from vggt_qwen3_b200 import ops
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    torch.manual_seed(0)
    m = PerceiverProjector(PerceiverConfig(256, 128, 2, 2, 512, dropout), in_dim=128, out_dim=192).cuda()
    m.train(dropout > 0)
    tokens = torch.randn(3, 72, 128, device="cuda")
    grad_out = torch.randn(3, 128, 192, device="cuda")
    res = {}
    try:
        for fused in (0, 1):
            ops.perceiver_attention_tune(fused)
            assert ops.perceiver_attention_supported(128, 72, 256, 2) == bool(fused)
            torch.manual_seed(7)                       # same dropout (seed, offset) for both runs
            m.zero_grad(set_to_none=True)
            t = tokens.clone().requires_grad_(True)
            out = m(t)
            out.backward(grad_out)
            res[fused] = (out.detach().clone(), t.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
    finally:
        ops.perceiver_attention_tune(1)
    (o0, g0, p0), (o1, g1, p1) = res[0], res[1]
    assert _rel(o1, o0) <= 5e-3 and _rel(g1, g0) <= 1e-2
    for n in p0:
        if n.endswith("in_proj_bias"):
            continue   # the key-bias third is identically zero in exact arithmetic: noise over noise
        assert _rel(p1[n], p0[n]) <= 1e-2, n
