"""GPU parity: CUDA aggregator (through the C ABI) vs the fp32 oracle restatement on the same seeded inputs."""
import pytest
import torch

from oracle.aggregator_ref import AggregatorConfig, aggregator_forward, init_aggregator_state
from tests.parity import assert_token_parity

pytestmark = pytest.mark.gpu


def _build(cfg: AggregatorConfig, state):
    from vggt_qwen3_b200.aggregator import Aggregator

    agg = Aggregator(img_size=cfg.img_size, patch_size=cfg.patch_size, embed_dim=cfg.embed_dim, depth=cfg.depth,
                     num_heads=cfg.num_heads, num_register_tokens=cfg.num_register_tokens, stem_depth=cfg.stem_depth)
    missing = agg.load_state_dict(state, strict=True)  # upstream key layout must match exactly
    return agg.cuda().eval()


CASES = [
    # (img, D, heads, depth, stem_depth, B, S, sensitised)
    (56, 128, 2, 2, 2, 2, 3, False),
    (56, 128, 2, 2, 2, 2, 3, True),
    (126, 256, 4, 3, 2, 1, 2, True),
    (154, 128, 2, 2, 1, 3, 1, True),   # 11x11 patches: P = 126, ragged vs every tile size
]


@pytest.mark.parametrize("img,D,heads,depth,stem_depth,B,S,sens", CASES)
def test_aggregator_matches_oracle(img, D, heads, depth, stem_depth, B, S, sens):
    cfg = AggregatorConfig(img_size=img, embed_dim=D, num_heads=heads, depth=depth, stem_depth=stem_depth)
    state = init_aggregator_state(cfg, seed=0, sensitised=sens)
    images = torch.rand(B, S, 3, img, img, generator=torch.Generator().manual_seed(1))
    ref_list, ref_idx = aggregator_forward(state, images, cfg, return_all=True)
    agg = _build(cfg, state)
    out_list, idx = agg(images.cuda(), return_all=True)
    assert idx == ref_idx == 5
    assert len(out_list) == len(ref_list) == depth
    for i, (o, r) in enumerate(zip(out_list, ref_list)):
        assert o.dtype == torch.bfloat16
        assert_token_parity(o, r, f"iteration {i}")
    # default call: only the last entry, same values
    last_only, _ = agg(images.cuda())
    assert len(last_only) == 1 and torch.equal(last_only[-1], out_list[-1])
    # encode path: first n tokens of the view-flattened last entry
    n = 17
    taken = agg.encode_tokens(images.cuda(), n)
    assert torch.equal(taken, out_list[-1].reshape(B, -1, 2 * D)[:, :n])


def test_layernorm_folding_matches_the_two_kernel_form():
    """The LayerNorm-folded schedule (5 launches per block) against the LayerNorm-kernel schedule (7): same tokens up
    to bf16 rounding, and both within the parity tolerance of the oracle (sensitised weights: non-trivial LN affine)."""
    cfg = AggregatorConfig(img_size=126, embed_dim=256, num_heads=4, depth=3, stem_depth=2)
    state = init_aggregator_state(cfg, seed=0, sensitised=True)
    images = torch.rand(2, 2, 3, 126, 126, generator=torch.Generator().manual_seed(1))
    ref_list, _ = aggregator_forward(state, images, cfg, return_all=False)
    agg = _build(cfg, state)
    from vggt_qwen3_b200 import _lib

    outs = {}
    for fold in (True, False):
        agg.fold_layernorm = fold
        agg.invalidate_packed()
        n0 = _lib.lib().vq_launch_count()
        outs[fold] = agg(images.cuda())[0][-1]
        torch.cuda.synchronize()
        outs[(fold, "launches")] = _lib.lib().vq_launch_count() - n0
        assert_token_parity(outs[fold], ref_list[-1], f"fold_layernorm={fold}")
    assert outs[(True, "launches")] < outs[(False, "launches")]
    assert_token_parity(outs[True], outs[False].float(), "folded vs two-kernel form", cos_min=0.9995, rel_max=1.5e-2)


def test_block_schedules_agree():
    """Default schedule (residual add deferred into the following LayerNorm pass, bf16 branch outputs) against the
    round-1 schedule (f32 read-modify-write in the GEMM epilogues): same tokens up to bf16 rounding of the branch
    outputs, both within the parity tolerance of the oracle; sensitised weights (LayerScale up to 1)."""
    cfg = AggregatorConfig(img_size=126, embed_dim=256, num_heads=4, depth=3, stem_depth=2)
    state = init_aggregator_state(cfg, seed=0, sensitised=True)
    images = torch.rand(2, 2, 3, 126, 126, generator=torch.Generator().manual_seed(1))
    ref_list, _ = aggregator_forward(state, images, cfg, return_all=True)
    agg = _build(cfg, state)
    outs = {}
    for deferred in (True, False):
        agg.deferred_residual = deferred
        lst, _ = agg(images.cuda(), return_all=True)
        for i, (o, r) in enumerate(zip(lst, ref_list)):
            assert_token_parity(o, r, f"deferred_residual={deferred} iteration {i}")
        outs[deferred] = lst[-1]
    assert_token_parity(outs[True], outs[False].float(), "deferred vs epilogue residual", cos_min=0.9995, rel_max=1.5e-2)


def test_uint8_views_equal_the_float_path_bit_for_bit():
    """uint8 views [B,S,H,W,3] straight into the patch kernel (x / 255 + normalise + im2col fused, 16-byte stores)
    give the same tokens, bit for bit, as ToTensor's float image through the f32 path."""
    cfg = AggregatorConfig(img_size=126, embed_dim=256, num_heads=4, depth=2, stem_depth=2)
    state = init_aggregator_state(cfg, seed=0, sensitised=True)
    agg = _build(cfg, state)
    views = torch.randint(0, 256, (2, 3, 126, 126, 3), generator=torch.Generator().manual_seed(4), dtype=torch.uint8)
    as_float = views.permute(0, 1, 4, 2, 3).float().div(255)          # torchvision ToTensor
    a = agg.encode_tokens(views.cuda(), None)
    b = agg.encode_tokens(as_float.cuda(), None)
    assert torch.equal(a, b)
    ref_list, _ = aggregator_forward(state, as_float, cfg, return_all=False)
    assert_token_parity(a.view(2, 3, -1, 512), ref_list[-1], "uint8 views")
    # odd base alignment of the byte strip (a view into a larger buffer)
    big = torch.zeros(views.numel() + 7, dtype=torch.uint8, device="cuda")
    for off in (1, 2, 3):
        v = big[off:off + views.numel()].view(views.shape)
        v.copy_(views)
        assert torch.equal(agg.encode_tokens(v, None), b)


@pytest.mark.parametrize("H,W,frames", [(518, 518, 3), (448, 448, 2), (126, 140, 1)])
def test_uint8_patch_matrix_equals_the_float_one(H, W, frames):
    """vq_im2col_patches_u8 (patch 14: raw byte strip in smem, 16-byte loads, table normalisation) against
    vq_im2col_patches on ToTensor's float image: the same bf16 patch matrix bit for bit, for every alignment of the
    view inside its buffer, including buffers that END with the last strip (no readable byte behind it)."""
    from vggt_qwen3_b200 import _lib

    lib = _lib.lib()
    Kp, gh, gw = 592, H // 14, W // 14
    views = torch.randint(0, 256, (frames, H, W, 3), generator=torch.Generator().manual_seed(H + frames), dtype=torch.uint8)
    img = views.permute(0, 3, 1, 2).float().div(255).contiguous().cuda()
    ref = torch.full((frames * gh * gw, Kp), 7.0, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.vq_im2col_patches(img.data_ptr(), ref.data_ptr(), frames, H, W, 14, Kp, st), "im2col")
    n = views.numel()
    for off in (0, 1, 5, 8, 15):
        big = torch.zeros(off + n, dtype=torch.uint8, device="cuda")      # the views end exactly at the buffer's end
        big[off:].copy_(views.flatten())
        out = torch.full_like(ref, -3.0)
        _lib.check(lib.vq_im2col_patches_u8(big.data_ptr() + off, out.data_ptr(), frames, H, W, 14, Kp, st), "im2col_u8")
        assert torch.equal(out, ref), f"offset {off}"
    assert (ref[:, 588:] == 0).all()
