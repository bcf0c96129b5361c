"""GPU parity of the kernels either side of the encode path, through the C ABI:
the inference splice against the oracle restatement of `_insert_vision_tokens` (bit-exact), the fused AdamW against
`torch.optim.AdamW` on the same gradients."""
import pytest
import torch

from oracle import perceiver_ref as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,n,hidden,where,dtype,mask_dtype", [
    (1, 40, 128, 2560, 7, torch.float32, torch.int64), (1, 12, 16, 64, 0, torch.bfloat16, torch.int64),
    (1, 12, 16, 64, 11, torch.float32, torch.bool), (2, 20, 8, 128, 25, torch.float32, torch.int32),
    (1, 9, 4, 32, None, torch.float32, torch.int64),
])
def test_splice_insert_matches_reference_restatement(B, L, n, hidden, where, dtype, mask_dtype):
    from vggt_qwen3_b200 import ops

    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 50, (B, L), generator=g)
    if where is not None:
        ids.view(-1)[where] = 777
        if where + 3 < B * L:
            ids.view(-1)[where + 3] = 777      # a later <image> token must be ignored
    emb = torch.randn(B, L, hidden, generator=g).to(dtype)
    mask = (torch.rand(B, L, generator=g) > 0.2).to(mask_dtype)
    vis = torch.randn(B, n, hidden, generator=g).to(dtype)
    out, out_mask = ops.splice_insert(emb.cuda(), mask.cuda(), ids.cuda(), vis.cuda(), 777)
    if where is None:
        ref, ref_mask = emb, mask
    else:
        # the reference slices every sample at the position of the first hit (qa_inference.py:129-135)
        pos = where % L
        ref = torch.cat([emb[:, :pos], vis, emb[:, pos + 1:]], dim=1)
        ref_mask = torch.cat([mask[:, :pos], torch.ones(B, n, dtype=mask_dtype), mask[:, pos + 1:]], dim=1)
        if B == 1:
            o_ref, o_mask = P.splice_insert(emb, mask, ids, vis, 777)
            assert torch.equal(o_ref, ref) and torch.equal(o_mask, ref_mask)
    assert out.shape == ref.shape and out_mask.shape == ref_mask.shape
    assert torch.equal(out.cpu(), ref) and torch.equal(out_mask.cpu(), ref_mask)


def test_fused_adamw_matches_torch_adamw():
    from vggt_qwen3_b200.optim import FusedAdamW

    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(257, 64), (4096,), (128, 512)]
    ps_ref = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone()) for p in ps_ref]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    ref = torch.optim.AdamW(ps_ref, **kw)
    opt = FusedAdamW(ps, **kw)
    sched_ref = torch.optim.lr_scheduler.LambdaLR(ref, lambda s: 1.0 / (1 + s))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    for _ in range(5):
        for a, b in zip(ps_ref, ps):
            gr = torch.randn(a.shape, device="cuda", generator=g)
            a.grad = gr.clone()
            b.grad = gr
        ref.step(); opt.step(); sched_ref.step(); sched.step()
    for a, b in zip(ps_ref, ps):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (a - b).abs().max()
        sa, sb = ref.state[a], opt.state[b]
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_adamw_step_bf16_gradients_scale_and_compute_copy():
    """The all-reduced gradient arrives in bf16 and unscaled (sum over ranks): grad_scale = 1/world folds the mean into
    the update, and the bf16 compute copy of the new weights is written in the same pass."""
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    n = 1 << 16
    p = torch.randn(n, device="cuda", generator=g)
    ref_p = torch.nn.Parameter(p.clone())
    ref = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    copy = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        gr = (torch.randn(n, device="cuda", generator=g) * 4).bfloat16()
        ref_p.grad = gr.float() * 0.25
        ref.step()
        ops.adamw_step(p, m, v, gr, step, lr=1e-2, weight_decay=0.1, grad_scale=0.25, param_bf16=copy)
    assert torch.allclose(p, ref_p.detach(), rtol=2e-5, atol=2e-6)
    assert torch.equal(copy, p.bfloat16())


def test_fused_adamw_refreshes_projector_compute_copies():
    """After a fused step the projector's next forward must use the UPDATED weights: the bf16 matrix copies are written
    by vq_adamw_step in the same pass, the f32 vector copies are re-copied lazily.  Checked against a fresh module
    loaded from the stepped module's state_dict (whose compute copies are rebuilt from the f32 masters)."""
    from vggt_qwen3_b200.optim import FusedAdamW
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    cfg = PerceiverConfig(256, 64, 2, 2, 512, 0.0)
    torch.manual_seed(0)
    b = PerceiverProjector(cfg, in_dim=256, out_dim=192).cuda().eval()
    before = {k: v.clone() for k, v in b.state_dict().items()}
    x = torch.randn(2, 64, 256, device="cuda").bfloat16()
    go = torch.randn(2, 64, 192, device="cuda")
    y0 = b(x).detach().float()
    ob = FusedAdamW(b.parameters(), lr=1e-2, weight_decay=0.01).attach_projector(b)
    for _ in range(2):
        ob.zero_grad(set_to_none=True)
        b(x).backward(go)
        ob.step()
    yb = b(x).detach().float()
    fresh = PerceiverProjector(cfg, in_dim=256, out_dim=192).cuda().eval()
    fresh.load_state_dict(b.state_dict())
    yf = fresh(x).detach().float()
    assert torch.equal(yb, yf), "stale compute copies after the fused optimizer step"
    assert (yb - y0).abs().max() > 1e-3, "the step did not change the output"
    changed = [k for k, v in b.state_dict().items() if not torch.equal(v, before[k])]
    assert len(changed) == len(before), "every parameter (matrices AND vectors) must have been updated"


@pytest.mark.parametrize("n,n_srcs,gdtype,wdtype", [
    (1 << 16, 1, torch.float32, torch.bfloat16), (4099, 3, torch.bfloat16, torch.bfloat16),
    (128 * 1000, 8, torch.bfloat16, torch.float32), (7, 2, torch.float32, torch.float32)])
def test_adamw_reduce_step_sums_sources_in_fp32(n, n_srcs, gdtype, wdtype):
    """vq_adamw_reduce_step: the gradient is the fp32 SUM of the sources (never rounded back to bf16 between adds),
    AdamW on the f32 master slice, compute copy written in the same pass; any n (scalar tail)."""
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n + n_srcs)
    p = torch.randn(n, device="cuda", generator=g)
    ref_p = torch.nn.Parameter(p.clone())
    ref = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    w = torch.empty(n, device="cuda", dtype=wdtype)
    for step in range(1, 4):
        srcs = [(torch.randn(n, device="cuda", generator=g) * (10.0 ** (i % 3))).to(gdtype) for i in range(n_srcs)]
        ref_p.grad = torch.stack([s.double() for s in srcs]).sum(0).float() / n_srcs
        ref.step()
        ops.adamw_reduce_step(p, m, v, srcs, w, step, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.1, grad_scale=1.0 / n_srcs)
    assert torch.allclose(p, ref_p.detach(), rtol=3e-5, atol=3e-6), (p - ref_p.detach()).abs().max()
    assert torch.equal(w, p.to(wdtype))
    with pytest.raises(RuntimeError):
        ops.adamw_reduce_step(p, m, v, srcs, w, 4, lr=1e-2, grad_scale=0.0)


@pytest.mark.parametrize("accumulate", [1, 2])
def test_sharded_adamw_single_gpu_training_steps(accumulate):
    """ShardedAdamW at world size 1 on the real kernels: bf16 wire gradients written by the wgrad GEMMs (or the f32
    accumulation buffer cast at the boundary), flat single-launch AdamW, compute copies updated in place.  Reference:
    the fp32 oracle's gradients through torch.optim.AdamW.  Adam normalises the gradient, so bf16 gradient error moves
    a first step by at most a few per cent of its size."""
    from vggt_qwen3_b200.optim import ShardedAdamW
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    shape = P.PerceiverShape(256, 64, 2, 2, 512, 0.0, 128, 192)
    state = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
    state["latents"] = state["latents"] * 25.0
    m = PerceiverProjector(PerceiverConfig(256, 64, 2, 2, 512, 0.0), in_dim=128, out_dim=192)
    m.load_state_dict(state, strict=True)
    m = m.cuda().eval()
    opt = ShardedAdamW(m, lr=1e-3, weight_decay=0.01, accumulate_steps=accumulate)
    ref_state = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    ref_opt = torch.optim.AdamW(list(ref_state.values()), lr=1e-3, weight_decay=0.01)
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        ref_opt.zero_grad()
        opt.zero_grad()
        for micro in range(accumulate):
            tokens = torch.randn(2, 64, shape.in_dim, generator=g)
            grad_out = torch.randn(2, shape.num_latents, shape.out_dim, generator=g)
            out = P.perceiver_forward(ref_state, tokens, shape.num_heads)
            out.backward(grad_out / accumulate)
            y = m(tokens.cuda())
            y.backward(grad_out.cuda())
            assert all(p.grad is None for p in m.parameters()) or accumulate > 1
        ref_opt.step()
        opt.step()
    opt.sync_parameters()
    for name, p in m.named_parameters():
        r, p0, got = ref_state[name].detach(), state[name], p.detach().cpu()
        if name.endswith("in_proj_bias"):
            # the gradient of the KEY bias is identically zero (softmax is invariant to a shift common to all keys); what
            # is left is rounding noise, which Adam normalises to +-lr steps of arbitrary sign: not comparable
            E = r.numel() // 3
            keep = torch.cat([torch.arange(0, E), torch.arange(2 * E, 3 * E)])
            r, p0, got = r[keep], p0[keep], got[keep]
        moved = (r - p0).norm()
        err = (got - r).norm()
        assert err <= 0.12 * moved + 1e-7, f"{name}: update error {err:.3e} vs update size {moved:.3e}"
    # the next forward reads the updated compute copies: same output as a fresh module loaded from the synced parameters
    x = torch.randn(2, 64, shape.in_dim, generator=g).cuda()
    fresh = PerceiverProjector(PerceiverConfig(256, 64, 2, 2, 512, 0.0), in_dim=128, out_dim=192).cuda().eval()
    fresh.load_state_dict(m.state_dict())
    assert torch.equal(m(x).detach(), fresh(x).detach())
    # load_state_dict on the module is picked up by the optimizer (masters + compute copies reloaded)
    m.load_state_dict(state)
    y_loaded = m(x).detach()
    fresh.load_state_dict(state)
    assert torch.equal(y_loaded, fresh(x).detach())
