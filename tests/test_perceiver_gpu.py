"""GPU parity of the Perceiver projector (forward + backward through the C ABI) against the reference's golden
vectors and against the oracle restatement on seeded inputs."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import perceiver_ref as P
from tests.parity import assert_token_parity

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"

# bf16 tensor-core compute vs fp32/fp64 reference: per-parameter gradient tolerance
GRAD_REL_L2 = 3e-2
GRAD_COS = 0.999


def _build(shape: P.PerceiverShape, state):
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    cfg = PerceiverConfig(latent_dim=shape.latent_dim, num_latents=shape.num_latents, num_heads=shape.num_heads,
                          num_layers=shape.num_layers, ffn_dim=shape.ffn_dim, dropout=shape.dropout)
    m = PerceiverProjector(cfg, in_dim=shape.in_dim, out_dim=shape.out_dim)
    m.load_state_dict(state, strict=True)  # reference key layout
    return m.cuda().eval()


def _check_grads(model, ref_grads):
    for name, p in model.named_parameters():
        g, r = p.grad.detach().double().cpu().flatten(), ref_grads[name].double().flatten()
        assert torch.isfinite(g).all(), name
        rel = (g - r).norm() / r.norm().clamp_min(1e-30)
        cos = torch.nn.functional.cosine_similarity(g, r, dim=0)
        assert rel <= GRAD_REL_L2, f"grad {name}: rel L2 err {rel:.3e}"
        assert cos >= GRAD_COS, f"grad {name}: cosine {cos:.6f}"


def test_reference_golden_forward_backward():
    z = np.load(GOLD / "perceiver_ref_small.npz")
    E, Lq, H, L, F, in_dim, out_dim = (int(v) for v in z["cfg"])
    state = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    m = _build(P.PerceiverShape(E, Lq, H, L, F, 0.1, in_dim, out_dim), state)
    tokens = torch.from_numpy(z["tokens"]).cuda().requires_grad_(True)
    out = m(tokens)
    assert out.dtype == torch.float32 and out.shape == z["out"].shape
    assert_token_parity(out, torch.from_numpy(z["out"]), "perceiver forward vs reference golden")
    out.backward(torch.from_numpy(z["grad_out"]).cuda())
    _check_grads(m, grads)
    gt, rt = tokens.grad.double().cpu().flatten(), torch.from_numpy(z["grad_tokens"]).double().flatten()
    assert (gt - rt).norm() / rt.norm() <= GRAD_REL_L2


@pytest.mark.parametrize("shape,B,T", [
    (P.PerceiverShape(512, 128, 4, 2, 1024, 0.0, 256, 320), 3, 128),
    (P.PerceiverShape(256, 64, 2, 3, 512, 0.0, 128, 192), 2, 72),     # T not a multiple of 64: padded keys
    (P.PerceiverShape(1024, 128, 8, 1, 2048, 0.0, 512, 640), 1, 64),  # head_dim 128, default num_vis_tokens
])
def test_matches_oracle(shape, B, T):
    state = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
    state["latents"] = state["latents"] * 25.0
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(B, T, shape.in_dim, generator=g)
    grad_out = torch.randn(B, shape.num_latents, shape.out_dim, generator=g)
    ref_out, ref_grads, ref_gt = P.perceiver_forward_backward(state, tokens, grad_out, shape.num_heads)
    m = _build(shape, state)
    tok = tokens.cuda().requires_grad_(True)
    out = m(tok)
    assert_token_parity(out, ref_out, "perceiver forward vs oracle")
    out.backward(grad_out.cuda())
    _check_grads(m, ref_grads)
    assert (tok.grad.double().cpu() - ref_gt.double()).norm() / ref_gt.double().norm() <= GRAD_REL_L2
    # second backward pass accumulates through autograd like any nn.Module
    out2 = m(tok)
    out2.backward(grad_out.cuda())
    for name, p in m.named_parameters():
        r = 2 * ref_grads[name].double().flatten()
        assert (p.grad.double().cpu().flatten() - r).norm() / r.norm() <= GRAD_REL_L2, name


def test_module_contract():
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    cfg = PerceiverConfig(latent_dim=128, num_latents=64, num_heads=2, num_layers=1, ffn_dim=256, dropout=0.1)
    m = PerceiverProjector(cfg, in_dim=64, out_dim=96)
    keys = set(m.state_dict().keys())
    assert keys == {
        "latents", "in_proj.weight", "in_proj.bias", "out_proj.weight", "out_proj.bias",
        "layers.0.self_attn.in_proj_weight", "layers.0.self_attn.in_proj_bias",
        "layers.0.self_attn.out_proj.weight", "layers.0.self_attn.out_proj.bias",
        "layers.0.mlp.0.weight", "layers.0.mlp.0.bias", "layers.0.mlp.3.weight", "layers.0.mlp.3.bias",
        "layers.0.norm1.weight", "layers.0.norm1.bias", "layers.0.norm2.weight", "layers.0.norm2.bias"}
    m = m.cuda()
    x = torch.randn(2, 64, 64, device="cuda")
    with pytest.raises(NotImplementedError):
        m.train()(x)          # dropout > 0 in training mode is refused, not silently dropped
    with pytest.raises(RuntimeError):
        m.eval()(x.cpu())     # no CPU fallback
    y = m.eval()(x)
    assert y.shape == (2, 64, 96) and y.requires_grad
    with torch.no_grad():
        assert not m(x).requires_grad
    # optimizer-style in-place update is picked up by the bf16 compute copies
    with torch.no_grad():
        m.out_proj.bias.add_(1.0)
    y2 = m(x)
    assert torch.allclose(y2 - y, torch.ones_like(y), atol=1e-2)
    # bf16 parameters (DeepSpeed bf16 policy) -> bf16 output
    mb = m.to(torch.bfloat16)
    assert mb(x.bfloat16()).dtype == torch.bfloat16
