"""GPU parity of vq_attention_fwd (through the C ABI) against a plain PyTorch fp32 reference of the same op:
softmax(q k^T) v with q pre-scaled by log2(e)/sqrt(d) (base-2 softmax), ragged and masked-tail sequence lengths,
both CTA shapes (one / two 128-row query tiles per CTA) and every FMA-pipe exp2 share.

Tolerance (north star): per-row cosine >= 0.999 and max |a-b| / max|b| per row <= 2e-2 for bf16 against fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(qkv, n_seq, L, heads):
    HD = heads * 64
    q, k, v = (qkv[:, i * HD:(i + 1) * HD].float().view(n_seq, L, heads, 64).transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * math.log(2.0)
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(n_seq * L, HD)


def _check(out, ref):
    a, b = out.float(), ref.float()
    a = a.view(-1, 64)
    b = b.view(-1, 64)
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
    rel = ((a - b).abs().amax(-1) / b.abs().amax(-1).clamp_min(1e-20)).max().item()
    assert cos >= 0.999 and rel <= 2e-2, f"cosine {cos:.6f}, max rel err {rel:.3e}"


@pytest.fixture(autouse=True)
def _restore_variant():
    from vggt_qwen3_b200 import ops

    yield
    ops.attention_tune(0, 4)


@pytest.mark.parametrize("qt", [1, 2, 3, 4])
@pytest.mark.parametrize("n_seq,L,heads,qscale", [
    (1, 128, 1, 2.0), (2, 64, 2, 2.0), (2, 65, 2, 2.0), (3, 50, 1, 2.0), (2, 129, 1, 2.0), (2, 200, 2, 2.0),
    (2, 300, 3, 2.0), (2, 1374, 2, 12.0), (1, 2000, 2, 0.05), (1, 2 * 1374, 4, 2.0), (1, 1, 1, 2.0), (1, 257, 2, 4.0),
])
def test_attention_matches_fp32_reference(qt, n_seq, L, heads, qscale):
    from vggt_qwen3_b200 import ops

    ops.attention_tune(qt, 2)
    g = torch.Generator(device="cuda").manual_seed(2 + L)
    HD = heads * 64
    qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
    qkv[:, :HD] *= qscale   # sharp (12) and flat (0.05) softmax rows exercise the lazy rescale and the exp2 range
    qkv = qkv.bfloat16()
    out = ops.attention_fwd(qkv, n_seq, L, heads)
    torch.cuda.synchronize()
    _check(out, _reference(qkv, n_seq, L, heads))


@pytest.mark.parametrize("emu", [0, 1, 2, 3, 4, 5, 6])
def test_attention_exp2_emulation_share(emu):
    """Every split between MUFU.EX2 and the FMA-pipe polynomial stays within the tolerance."""
    from vggt_qwen3_b200 import ops

    ops.attention_tune(0, emu)
    g = torch.Generator(device="cuda").manual_seed(7)
    n_seq, L, heads = 2, 700, 2
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g)
    qkv[:, :heads * 64] *= 6.0
    qkv = qkv.bfloat16()
    out = ops.attention_fwd(qkv, n_seq, L, heads)
    torch.cuda.synchronize()
    _check(out, _reference(qkv, n_seq, L, heads))


def test_attention_growing_maximum():
    """Keys sorted so that the row maximum keeps growing tile after tile: every lazy-rescale branch is taken."""
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    n_seq, L, heads = 1, 1024, 1
    q = torch.randn(L, 64, device="cuda", generator=g).abs() * 2.0
    k = torch.randn(L, 64, device="cuda", generator=g).abs() * torch.linspace(0.1, 4.0, L, device="cuda")[:, None]
    v = torch.randn(L, 64, device="cuda", generator=g)
    qkv = torch.cat([q, k, v], dim=1).bfloat16()
    for qt in (1, 2, 3, 4):
        ops.attention_tune(qt, 2)
        out = ops.attention_fwd(qkv, n_seq, L, heads)
        torch.cuda.synchronize()
        _check(out, _reference(qkv, n_seq, L, heads))


def test_attention_sequences_are_independent():
    """Frame attention = n_seq independent sequences over one buffer: changing sequence 1 leaves sequence 0 bit-identical."""
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    n_seq, L, heads = 3, 333, 2
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g).bfloat16()
    a = ops.attention_fwd(qkv, n_seq, L, heads).clone()
    qkv2 = qkv.clone()
    qkv2[L:2 * L] = torch.randn(L, 3 * heads * 64, device="cuda", generator=g).bfloat16()
    b = ops.attention_fwd(qkv2, n_seq, L, heads)
    torch.cuda.synchronize()
    assert torch.equal(a[:L], b[:L]) and torch.equal(a[2 * L:], b[2 * L:])
    assert not torch.equal(a[L:2 * L], b[L:2 * L])


@pytest.mark.parametrize("emu", [0, 3, 6])
@pytest.mark.parametrize("n_seq,L,heads", [(1, 128, 1), (2, 64, 2), (2, 65, 2), (3, 50, 1), (2, 129, 1), (2, 200, 2),
                                           (2, 1374, 2), (1, 1, 1), (1, 257, 2), (1, 2 * 1374, 4), (1, 191, 1), (1, 193, 1)])
def test_attention_bounded_variant(n_seq, L, heads, emu):
    """vq_attention_fwd_bounded: scores proven <= 64 in magnitude (here: unit-norm-scaled q, k as after the per-head
    LayerNorm of the aggregator blocks) -> no range check, two threads per query row; ragged tails in either key half."""
    from vggt_qwen3_b200 import ops

    ops.attention_tune_bounded(1, emu)
    g = torch.Generator(device="cuda").manual_seed(5 + L)
    HD = heads * 64
    qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
    # rows of q and k normalised per head to norm 8 (what LayerNorm(64) produces), q additionally scaled by log2(e)/8:
    # |s| <= 8 * 8 * 0.18 = 11.6
    for blk, scale in ((0, math.log2(math.e) / 8.0), (1, 1.0)):
        v = qkv[:, blk * HD:(blk + 1) * HD].view(-1, heads, 64)
        v.copy_(v / v.norm(dim=-1, keepdim=True) * 8.0 * scale)
    qkv = qkv.bfloat16()
    out = ops.attention_fwd_bounded(qkv, n_seq, L, heads, 11.6 * 1.02)
    torch.cuda.synchronize()
    _check(out, _reference(qkv, n_seq, L, heads))
    # the same call with the variant switched off (or without a bound) is the checked kernel: same tolerance
    ops.attention_tune_bounded(0, -1)
    _check(ops.attention_fwd_bounded(qkv, n_seq, L, heads, 11.6), _reference(qkv, n_seq, L, heads))
    ops.attention_tune_bounded(1, 3)
    _check(ops.attention_fwd_bounded(qkv, n_seq, L, heads, 0.0), _reference(qkv, n_seq, L, heads))


def test_attention_bounded_at_the_bound():
    """Scores that actually reach +-60 (inside the promised 64): P = 2^s spans 2^-60 .. 2^60 without a reference."""
    from vggt_qwen3_b200 import ops

    ops.attention_tune_bounded(1, 3)
    g = torch.Generator(device="cuda").manual_seed(9)
    n_seq, L, heads = 1, 700, 1
    q = torch.randn(L, 64, device="cuda", generator=g)
    k = torch.randn(L, 64, device="cuda", generator=g)
    q = q / q.norm(dim=-1, keepdim=True) * 7.7
    k = k / k.norm(dim=-1, keepdim=True) * 7.7
    k[::50] = q[::50] / 7.7 * 7.7          # aligned pairs: s = 59.3
    v = torch.randn(L, 64, device="cuda", generator=g)
    qkv = torch.cat([q, k, v], dim=1).bfloat16()
    s_max = (qkv[:, :64].float() @ qkv[:, 64:128].float().t()).abs().max().item()
    assert 55.0 < s_max <= 64.0
    out = ops.attention_fwd_bounded(qkv, n_seq, L, heads, 64.0)
    torch.cuda.synchronize()
    _check(out, _reference(qkv, n_seq, L, heads))
