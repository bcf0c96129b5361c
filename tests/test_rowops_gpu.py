"""GPU unit parity of the HBM-bound row kernels against plain PyTorch fp32 references."""
import math

import pytest
import torch

from oracle import perceiver_ref as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D,dtype", [(1374, 1024, torch.float32), (77, 4096, torch.bfloat16), (5, 128, torch.float32)])
def test_layernorm_forward(rows, D, dtype):
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(rows, D, device="cuda", generator=g) * 3 + 1).to(dtype)
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), w, b, 1e-5)
    y32 = ops.layernorm(x, w, b, 1e-5, out_dtype=torch.float32)
    assert (y32 - ref).abs().max() <= 2e-5 * ref.abs().max()
    y16, mean, rstd = ops.layernorm(x, w, b, 1e-5, out_dtype=torch.bfloat16, return_stats=True)
    assert (y16.float() - ref).abs().max() <= 1e-2 * ref.abs().max()
    assert torch.allclose(mean, x.float().mean(-1), atol=1e-4)
    assert torch.allclose(rstd, 1 / torch.sqrt(x.float().var(-1, unbiased=False) + 1e-5), rtol=1e-4)


def test_layernorm_backward():
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    rows, D = 300, 512
    z = (torch.randn(rows, D, device="cuda", generator=g) * 2).bfloat16()
    dy = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.zeros(D, device="cuda")
    zr = z.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(zr, (D,), wr, br, 1e-5).backward(dy.float())
    _, mean, rstd = ops.layernorm(z, w, b, 1e-5, return_stats=True)
    dz, dgamma, dbeta = ops.layernorm_bwd(dy, z, mean, rstd, w)
    assert (dz.float() - zr.grad).abs().max() <= 1e-2 * zr.grad.abs().max()
    assert (dgamma - wr.grad).abs().max() <= 1e-3 * wr.grad.abs().max()
    assert (dbeta - br.grad).abs().max() <= 1e-3 * br.grad.abs().max()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_splice_overwrite_matches_oracle(dtype):
    from vggt_qwen3_b200 import ops

    g = torch.Generator().manual_seed(0)
    B, L, H, n = 4, 300, 2560, 136
    ids = torch.randint(0, 1000, (B, L), generator=g)
    ids[0, 3] = 151_665
    ids[1, 10] = 151_665
    ids[1, 100] = 151_665      # second span overlaps the first: later wins
    ids[2, L - n] = 151_665    # fits exactly
    emb = torch.randn(B, L, H, generator=g).to(dtype)
    feats = torch.randn(B, n, H, generator=g).to(dtype)
    ref = P.splice_overwrite(emb, ids, feats, 151_665)
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = ops.splice_overwrite(emb.cuda().clone(), ids.cuda(), feats.cuda(), 151_665, overflow)
    assert torch.equal(out.cpu(), ref)          # bit exact: pure data movement
    assert int(overflow) == 0
    ids[3, L - 5] = 151_665                      # does not fit: counted, clipped
    out = ops.splice_overwrite(emb.cuda().clone(), ids.cuda(), feats.cuda(), 151_665, overflow)
    assert int(overflow) == 1
    assert torch.equal(out[3, L - 5:].cpu(), feats[3, :5])
    # no <image> token at all: untouched
    none = ops.splice_overwrite(emb.cuda().clone(), torch.zeros(B, L, dtype=torch.long, device="cuda"), feats.cuda(), 151_665)
    assert torch.equal(none.cpu(), emb)
