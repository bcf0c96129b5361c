"""GPU parity at BASELINE.json's own sizes (the configs the bench times): long-sequence attention with the kernel
variant `vq_attention_fwd` dispatches by itself (8-frame global attention L = 10 992, 32-frame L = 43 968), the
full-size aggregator (D = 1024, 24 + 48 blocks) on multi-view scenes at 518 px and at the 448 px the reference trains
at (configs/stage1_3d.yaml:19: 32 x 32 patches from the 37 x 37 position table), and the full-size Perceiver
(4096 / 16384, B = 8) forward + backward against a golden made by the reference's own class.

The fp32 oracle runs on the GPU here (tests/parity.py: fp32_oracle_on_cuda — TF32 off, math SDPA); one case repeats it
on the CPU to show both agree."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import aggregator_ref as A
from oracle import perceiver_ref as P
from tests.parity import assert_token_parity, fp32_oracle_on_cuda

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------------------------------------ attention
def _attention_reference(qkv, n_seq, L, heads, chunk=2048):
    """softmax_2(q k^T) v in fp32, query-chunked (q is pre-scaled by log2(e)/sqrt(d): base-2 softmax)."""
    HD = heads * 64
    out = torch.empty(n_seq * L, HD, device=qkv.device, dtype=torch.float32)
    for s in range(n_seq):
        rows = slice(s * L, (s + 1) * L)
        for h in range(heads):
            q = qkv[rows, h * 64:(h + 1) * 64].float()
            k = qkv[rows, HD + h * 64:HD + (h + 1) * 64].float()
            v = qkv[rows, 2 * HD + h * 64:2 * HD + (h + 1) * 64].float()
            for c in range(0, L, chunk):
                e = min(L, c + chunk)
                sc = (q[c:e] @ k.t()) * math.log(2.0)
                out[s * L + c:s * L + e, h * 64:(h + 1) * 64] = torch.softmax(sc, dim=-1) @ v
    return out


def _check_rows(out, ref):
    a, b = out.float().view(-1, 64), ref.float().view(-1, 64)
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
    rel = ((a - b).abs().amax(-1) / b.abs().amax(-1).clamp_min(1e-20)).max().item()
    assert cos >= 0.999 and rel <= 2e-2, f"cosine {cos:.6f}, max rel err {rel:.3e}"


@pytest.mark.parametrize("n_seq,L,heads", [(2, 8 * 1374, 2), (1, 32 * 1374, 1), (1, 16384 + 1374, 1)])
@pytest.mark.parametrize("kind", ["model_like", "outlier_rows"])
def test_attention_long_sequence_auto_dispatch(n_seq, L, heads, kind):
    """The variant chosen by vq_attention_fwd itself (wide kernel, one query tile per CTA up to 16 384 keys, two beyond):
    model-like scores (LayerNorm-ed q, k: |s| of a few units, the no-max-subtraction fast path) and rows that leave
    the fast path (|s| > 64 -> true row maximum, lazy rescale) in the middle of a long sequence."""
    from vggt_qwen3_b200 import ops

    ops.attention_tune(0, 4)
    with fp32_oracle_on_cuda():
        g = torch.Generator(device="cuda").manual_seed(L + heads)
        HD = heads * 64
        qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
        qkv[:, :HD] *= 0.9            # |s| = |q.k| ~ N(0, 0.9^2 * 64): a few units, max ~ 30
        if kind == "outlier_rows":
            hot = torch.randperm(n_seq * L, device="cuda", generator=g)[:257]
            qkv[hot, :HD] *= 9.0      # |s| up to ~ 250 for these rows: general path, rescales
            qkv[L // 2 + 77, HD:2 * HD] *= 6.0   # one key that dominates every row after half of the sequence
        qkv = qkv.bfloat16()
        out = ops.attention_fwd(qkv, n_seq, L, heads)
        torch.cuda.synchronize()
        _check_rows(out, _attention_reference(qkv, n_seq, L, heads))


@pytest.mark.parametrize("n_seq,L,heads", [(2, 8 * 1374, 2), (1, 32 * 1374, 1)])
def test_attention_long_sequence_bounded_variant(n_seq, L, heads):
    """The kernel the aggregator's frame / global blocks run (scores bounded through the QK LayerNorm) at the 8- and
    32-frame global sequence lengths."""
    from vggt_qwen3_b200 import ops

    ops.attention_tune_bounded(1, 3)
    with fp32_oracle_on_cuda():
        g = torch.Generator(device="cuda").manual_seed(L)
        HD = heads * 64
        qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
        for blk, scale in ((0, math.log2(math.e) / 8.0), (1, 1.0)):
            v = qkv[:, blk * HD:(blk + 1) * HD].view(-1, heads, 64)
            v.copy_(v / v.norm(dim=-1, keepdim=True) * 8.0 * scale * 1.5)     # |s| <= 26
        qkv = qkv.bfloat16()
        out = ops.attention_fwd_bounded(qkv, n_seq, L, heads, 27.0)
        torch.cuda.synchronize()
        _check_rows(out, _attention_reference(qkv, n_seq, L, heads))


# ------------------------------------------------------------------------------------------------ aggregator
def _state_to(state, device):
    return {k: v.to(device) for k, v in state.items()}


def _run_full_size(S, px, sensitised, B=1, seed=0):
    from vggt_qwen3_b200.aggregator import Aggregator

    cfg = A.AggregatorConfig()
    state = A.init_aggregator_state(cfg, seed=seed, sensitised=sensitised)
    images = torch.rand(B, S, 3, px, px, generator=torch.Generator().manual_seed(3))
    agg = Aggregator()
    agg.load_state_dict(state, strict=True)
    agg = agg.cuda().eval()
    lst, start = agg(images.cuda())
    torch.cuda.synchronize()
    del agg
    with fp32_oracle_on_cuda():
        ref_list, ref_start = A.aggregator_forward(_state_to(state, "cuda"), images.cuda(), cfg, return_all=False)
    P_ = 5 + (px // 14) ** 2
    assert start == ref_start == 5 and lst[-1].shape == ref_list[-1].shape == (B, S, P_, 2048)
    return lst[-1], ref_list[-1], state, images, cfg


@pytest.mark.parametrize("S,px,sensitised", [(2, 518, False), (8, 518, False), (8, 518, True), (2, 448, True)])
def test_full_size_multiview_aggregator(S, px, sensitised):
    """D = 1024, 24 DINOv2 + 24 x (frame, global) blocks on one S-frame scene: global attention over S * P keys inside
    the model (10 992 keys at S = 8), every token of the last iteration compared with the fp32 oracle.  `sensitised`
    = LayerScale ~ U(0.1, 1), live special tokens and QK-norm affine, so a block bug cannot hide under the 0.01 init."""
    out, ref, *_ = _run_full_size(S, px, sensitised)
    assert_token_parity(out, ref, f"full-size aggregator S={S} {px}px sensitised={sensitised}")


def test_gpu_oracle_equals_cpu_oracle():
    """The fp32 oracle gives the same answer on `cuda` (how the big cases above run it) as on the CPU."""
    cfg = A.AggregatorConfig(img_size=126, embed_dim=256, num_heads=4, depth=3, stem_depth=2)
    state = A.init_aggregator_state(cfg, seed=0, sensitised=True)
    images = torch.rand(2, 3, 3, 126, 126, generator=torch.Generator().manual_seed(1))
    cpu, _ = A.aggregator_forward(state, images, cfg, return_all=False)
    with fp32_oracle_on_cuda():
        gpu, _ = A.aggregator_forward(_state_to(state, "cuda"), images.cuda(), cfg, return_all=False)
    err = (gpu[-1].cpu() - cpu[-1]).abs().max() / cpu[-1].abs().max()
    assert err < 2e-5, err


def test_multiview_scene_changes_first_frame_tokens():
    """Frames 1..S-1 reach the 128 tokens encode_images keeps (all from frame 0) only through global attention: the
    S = 8 result must differ from the single-view result of the same first frame, and both match the oracle."""
    out8, ref8, state, images, cfg = _run_full_size(8, 518, True)
    from vggt_qwen3_b200.aggregator import Aggregator

    agg = Aggregator()
    agg.load_state_dict(state, strict=True)
    agg = agg.cuda().eval()
    one = agg.encode_tokens(images[:, :1].cuda(), 128)
    eight = out8.reshape(1, -1, 2048)[:, :128]
    assert (one.float() - eight.float()).abs().max() > 0.05
    assert_token_parity(eight, ref8.reshape(1, -1, 2048)[:, :128], "first 128 tokens of the 8-frame scene")


# ------------------------------------------------------------------------------------------------ Perceiver
def _probe(t: torch.Tensor, index: int, n_probes=4, n_sample=4096):
    """Mirror of tests/golden/make_golden_full.py:probe (norm, four seeded +-1 projections, strided sample)."""
    flat = t.detach().reshape(-1).double().cpu()
    out = [float(flat.norm())]
    for k in range(n_probes):
        g = torch.Generator().manual_seed(1000 * index + k)
        r = torch.randint(0, 2, (flat.numel(),), generator=g, dtype=torch.int8).double() * 2 - 1
        out.append(float(flat @ r))
    stride = max(1, flat.numel() // n_sample)
    return np.concatenate([np.array(out), flat[::stride][:n_sample].numpy()])


def _check_probe(got: np.ndarray, ref: np.ndarray, what: str, tol=5e-2):
    """bf16 tensor-core backward through six layers against the fp32 reference: relative L2 error <= 5e-2 and cosine
    >= 0.999 on the strided sample (a cosine of 0.999 is a relative error of 4.5e-2); the +-1 projections of the
    error are sums with standard deviation |error|, so they are bounded at three times that."""
    norm = ref[0]
    assert abs(got[0] - norm) <= tol * norm, f"{what}: norm {got[0]:.6e} vs {norm:.6e}"
    for k in range(1, 5):
        assert abs(got[k] - ref[k]) <= 3 * tol * norm, f"{what}: probe {k}: {got[k]:.6e} vs {ref[k]:.6e} (norm {norm:.3e})"
    a, b = torch.from_numpy(got[5:]), torch.from_numpy(ref[5:])
    rel, cos = 0.0, 1.0
    if b.norm() > 0:
        rel = ((a - b).norm() / b.norm()).item()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    _ERRORS.append(f"{what:50s} sample rel L2 {rel:.3e} cosine {cos:.6f} probe err/norm "
                   f"{max(abs(got[k] - ref[k]) for k in range(1, 5)) / max(norm, 1e-30):.3e}")
    assert rel <= tol and cos >= 0.999, f"{what}: sample rel L2 {rel:.3e}, cosine {cos:.6f}"


_ERRORS = []


def test_full_size_perceiver_forward_backward_vs_reference_class():
    """perceiver_small at full size on the c2 shape (B = 8 x 128 tokens; CTA-pair 256 x 256 tiles, MN-major wgrads
    with 1024 contraction rows) against the fp32 CPU run of the reference's own PerceiverProjector
    (tests/golden/perceiver_ref_full.npz, made by tests/golden/make_golden_full.py).  Weights and inputs are
    regenerated from the seeds: the module mirror draws its parameters exactly like the reference class."""
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    z = np.load(GOLD / "perceiver_ref_full.npz")
    torch.manual_seed(0)
    m = PerceiverProjector(PerceiverConfig(), in_dim=2048, out_dim=2560).eval()
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(8, 128, 2048, generator=g)
    grad_out = torch.randn(8, 128, 2560, generator=g)
    m = m.cuda()
    tok = tokens.cuda().requires_grad_(True)
    out = m(tok)
    out.backward(grad_out.cuda())
    torch.cuda.synchronize()
    ref_rows = torch.from_numpy(z["out_rows"])
    got_rows = out.detach()[:, ::16, :].reshape(-1, out.shape[-1])[:, ::4].float().cpu()
    assert_token_parity(got_rows, ref_rows, "full-size perceiver output rows")
    _check_probe(_probe(out, 0), z["out"], "output")
    _check_probe(_probe(tok.grad, 1), z["grad_tokens"], "grad tokens")
    try:
        for i, (n, p) in enumerate(m.named_parameters()):
            _check_probe(_probe(p.grad, 2 + i), z["grad/" + n], f"grad {n}")
    finally:   # per-tensor error table for profiles/ (only where the scratch directory exists)
        out_dir = Path(__file__).resolve().parent.parent / "gpurun_out"
        if out_dir.is_dir():
            (out_dir / "perceiver_full_errors.txt").write_text("\n".join(_ERRORS) + "\n")
