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


# ---------------------------------------------------------------------------------------------------- train mode
def _set_generator(seed: int, offset: int):
    torch.cuda.manual_seed(seed)
    torch.cuda.default_generators[torch.cuda.current_device()].set_offset(offset)


@pytest.mark.parametrize("rows,cols,ld,p,seed,offset,site", [
    (37, 72, 128, 0.1, 20240611, 12, 0), (5, 256, 256, 0.5, 2**63 + 5, 2**33 + 4, 7), (1024, 4096, 4096, 0.1, 3, 8, 21)])
def test_dropout_mask_matches_oracle(rows, cols, ld, p, seed, offset, site):
    from vggt_qwen3_b200 import ops

    got = ops.dropout_mask(rows, cols, ld, p, seed, offset, site).cpu().bool()
    assert torch.equal(got, P.dropout_keep_mask(rows, cols, ld, p, seed, offset, site))


def test_train_mode_matches_reference_class_golden():
    """`.train()` with dropout 0.1 — how the reference trains (src/train/train_sft.py:196) — against the golden made
    by the reference's own PerceiverProjector under .train() with the same counter-based masks: output, every
    parameter gradient and the token gradient, value for value."""
    z = np.load(GOLD / "perceiver_ref_train.npz")
    E, Lq, H, L, F, in_dim, out_dim = (int(v) for v in z["cfg"])
    p, seed, offset = float(z["dropout"][0]), int(z["dropout"][1]), int(z["dropout"][2])
    state = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    m = _build(P.PerceiverShape(E, Lq, H, L, F, p, in_dim, out_dim), state).train()
    tokens = torch.from_numpy(z["tokens"]).cuda().requires_grad_(True)
    _set_generator(seed, offset)
    out = m(tokens)
    assert torch.cuda.default_generators[torch.cuda.current_device()].get_offset() == offset + 4
    assert_token_parity(out, torch.from_numpy(z["out"]), "train-mode forward vs reference golden")
    out.backward(torch.from_numpy(z["grad_out"]).cuda())
    _check_grads(m, grads)
    gt, rt = tokens.grad.double().cpu().flatten(), torch.from_numpy(z["grad_tokens"]).double().flatten()
    assert (gt - rt).norm() / rt.norm() <= GRAD_REL_L2
    # eval mode on the same module is the dropout-free function again
    ev = m.eval()(tokens.detach())
    ref_ev = P.perceiver_forward(state, torch.from_numpy(z["tokens"]), H)
    assert_token_parity(ev, ref_ev, "eval after train")


def test_train_mode_full_width_matches_oracle():
    """Every tile shape of the dropout epilogues (CTA pairs, 256-wide tiles) and a padded key count."""
    shape = P.PerceiverShape(1024, 128, 8, 2, 4096, 0.1, 512, 640)
    B, T = 4, 72
    state = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
    state["latents"] = state["latents"] * 25.0
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(B, T, shape.in_dim, generator=g)
    grad_out = torch.randn(B, shape.num_latents, shape.out_dim, generator=g)
    seed, offset = 99, 40
    ref_out, ref_grads, ref_gt = P.perceiver_forward_backward(state, tokens, grad_out, shape.num_heads,
                                                              dropout=(0.1, seed, offset))
    m = _build(shape, state).train()
    tok = tokens.cuda().requires_grad_(True)
    _set_generator(seed, offset)
    out = m(tok)
    assert_token_parity(out, ref_out, "train-mode forward vs oracle")
    out.backward(grad_out.cuda())
    _check_grads(m, ref_grads)
    assert (tok.grad.double().cpu() - ref_gt.double()).norm() / ref_gt.double().norm() <= GRAD_REL_L2


def test_train_mode_semantics():
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    torch.manual_seed(0)
    x = torch.randn(2, 64, 64, device="cuda")
    m0 = PerceiverProjector(PerceiverConfig(128, 64, 2, 2, 256, 0.0), in_dim=64, out_dim=96).cuda()
    # p = 0: train mode is bit-identical to eval mode and does not touch the generator
    off = torch.cuda.default_generators[torch.cuda.current_device()].get_offset()
    assert torch.equal(m0.train()(x), m0.eval()(x))
    assert torch.cuda.default_generators[torch.cuda.current_device()].get_offset() == off
    m = PerceiverProjector(PerceiverConfig(128, 64, 2, 2, 256, 0.1), in_dim=64, out_dim=96).cuda()
    m.load_state_dict(m0.state_dict())
    assert torch.equal(m.eval()(x), m0(x))                  # eval ignores p
    m.train()
    torch.cuda.manual_seed(5)
    a, b = m(x), m(x)
    assert not torch.equal(a, b)                            # the generator offset advances between calls
    torch.cuda.manual_seed(5)
    assert torch.equal(m(x), a)                             # reproducible under the same seed
    assert not torch.equal(a, m.eval()(x))


def test_gradient_accumulation_in_place():
    """`p.grad` views alias the persistent flat buffer; a second backward with p.grad still set accumulates in the
    kernels (no autograd copy), `zero_grad(set_to_none=False)` keeps the aliasing."""
    shape = P.PerceiverShape(256, 64, 2, 2, 512, 0.0, 128, 192)
    state = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(2, 64, shape.in_dim, generator=g)
    grad_out = torch.randn(2, shape.num_latents, shape.out_dim, generator=g)
    _, ref_grads, _ = P.perceiver_forward_backward(state, tokens, grad_out, shape.num_heads)
    m = _build(shape, state)
    tok = tokens.cuda()
    m(tok).backward(grad_out.cuda())
    flat = m.grad_flat
    ptrs = [p.grad.data_ptr() for p in m._ordered_params()]
    assert all(flat.data_ptr() <= q < flat.data_ptr() + flat.numel() * 4 for q in ptrs)
    _check_grads(m, ref_grads)
    m(tok).backward(grad_out.cuda())                       # accumulate: 2x
    assert [p.grad.data_ptr() for p in m._ordered_params()] == ptrs and m.grad_flat is flat
    _check_grads(m, {k: 2 * v for k, v in ref_grads.items()})
    m.zero_grad(set_to_none=False)
    m(tok).backward(grad_out.cuda())
    assert [p.grad.data_ptr() for p in m._ordered_params()] == ptrs
    _check_grads(m, ref_grads)
    m.zero_grad(set_to_none=True)
    m(tok).backward(grad_out.cuda())
    assert [p.grad.data_ptr() for p in m._ordered_params()] == ptrs
    _check_grads(m, ref_grads)
