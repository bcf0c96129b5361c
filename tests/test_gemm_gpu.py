"""GPU parity of vq_gemm (through the C ABI) against plain PyTorch fp32 on the bf16-rounded operands: every fused
epilogue, both operand majors, ragged M / N / K, batched problems and the CTA-pair path."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def _operands(M, N, K, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    return g, A, B, A.float() @ B.float().t()


@pytest.mark.parametrize("M,N,K,a_major,b_major,f32", [
    (128, 128, 64, 0, 0, False), (128, 256, 256, 0, 0, False), (1374, 1024, 1024, 0, 0, False),
    (300, 1024, 592, 0, 0, False), (777, 512, 320, 0, 0, True), (1024, 1024, 512, 0, 1, False),
    (1024, 1024, 512, 1, 0, False), (4096, 2048, 1024, 1, 1, True), (5, 128, 64, 0, 0, True),
    (10992, 3072, 1024, 0, 0, False),
])
def test_gemm_linear(M, N, K, a_major, b_major, f32):
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm(A.t().contiguous() if a_major else A, B.t().contiguous() if b_major else B, a_major=a_major,
                   b_major=b_major, bias=bias, out_dtype=torch.float32 if f32 else torch.bfloat16)
    torch.cuda.synchronize()
    assert _rel(out, ref + bias) < (2e-3 if f32 else 1e-2)


@pytest.mark.parametrize("M,N,K", [(1374, 4096, 1024), (130, 256, 128)])
def test_gemm_gelu_and_saved_preactivation(M, N, K):
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, aux=aux)
    torch.cuda.synchronize()
    pre = ref + bias
    assert _rel(out, torch.nn.functional.gelu(pre)) < 1e-2
    assert _rel(aux, pre) < 1e-2


@pytest.mark.parametrize("M,N,K,f32", [(1374, 1024, 4096, True), (1024, 4096, 4096, False), (10992, 1024, 1024, True)])
def test_gemm_layerscale_residual(M, N, K, f32):
    """proj / fc2 epilogue: C = add + gamma * (acc + bias), in place on the f32 residual stream."""
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    if not f32:
        res = res.bfloat16()
    expect = res.float() + gamma * (ref + bias)
    out = ops.gemm(A, B, bias=bias, gamma=gamma, add=res, out_dtype=res.dtype)
    torch.cuda.synchronize()
    assert _rel(out, expect) < (2e-3 if f32 else 1e-2)
    if f32:  # in place: C aliases add
        x = res.clone()
        ops.gemm(A, B, bias=bias, gamma=gamma, add=x, out=x, out_dtype=torch.float32)
        torch.cuda.synchronize()
        assert torch.equal(x, out)


def test_gemm_gelu_backward():
    from vggt_qwen3_b200 import ops

    M, N, K = 1024, 2048, 512
    g, A, B, ref = _operands(M, N, K)
    pre = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    x = pre.float()
    cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    out = ops.gemm(A, B, epilogue=ops.EPI_GELU_BWD, aux=pre)
    torch.cuda.synchronize()
    assert _rel(out, ref * (cdf + x * pdf)) < 1e-2


@pytest.mark.parametrize("frames,grid,heads", [(3, 6, 2), (2, 37, 4), (5, 9, 16)])
def test_gemm_qkv_qknorm_rope_epilogue(frames, grid, heads):
    """qkv projection with the fused per-head QK LayerNorm(64) + 2-D RoPE + q pre-scale ([UPSTREAM]
    vggt/layers/attention.py q_norm / k_norm, vggt/layers/rope.py) against the ORACLE's own functions
    (oracle.aggregator_ref.rope_tables / rope_2d / token_positions) evaluated in fp32 on the same bf16-rounded operands."""
    from oracle import aggregator_ref as A
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    D, P = heads * 64, 5 + grid * grid
    M = frames * P
    x = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    bias = torch.randn(3 * D, device="cuda", generator=g) * 0.1
    qn_w, kn_w = (torch.rand(64, device="cuda", generator=g) + 0.5 for _ in range(2))
    qn_b, kn_b = (torch.randn(64, device="cuda", generator=g) * 0.1 for _ in range(2))
    cos, sin = (t.cuda().contiguous() for t in A.rope_tables(grid, 64, 100.0))
    q_scale = math.log2(math.e) / 8.0
    extra = dict(tokens_per_frame=P, patch_start=5, grid_w=grid, apply_qknorm_rope=1, qn_w=qn_w, qn_b=qn_b, kn_w=kn_w,
                 kn_b=kn_b, qk_eps=1e-5, q_scale=q_scale, rope_cos=cos, rope_sin=sin, rope_positions=grid + 1)
    out = ops.gemm(x, w, bias=bias, epilogue=ops.EPI_QKV_ROPE, extra=extra)
    torch.cuda.synchronize()
    qkv = (x.float() @ w.float().t() + bias).view(frames, P, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q = torch.nn.functional.layer_norm(qkv[0], (64,), qn_w, qn_b, 1e-5)
    k = torch.nn.functional.layer_norm(qkv[1], (64,), kn_w, kn_b, 1e-5)
    pos = A.token_positions(grid, grid, 5).cuda()[None].expand(frames, P, 2)
    q = A.rope_2d(q, pos, cos, sin) * q_scale
    k = A.rope_2d(k, pos, cos, sin)
    ref = torch.stack([q, k, qkv[2]], 0).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("M,N,K,f32", [(300, 512, 256, False), (1024, 4096, 1024, False), (260, 256, 128, True)])
def test_gemm_dropout_epilogues(M, N, K, f32):
    """Train-mode dropout fused into the LINEAR (+ residual), GELU and GELU_BWD epilogues: the mask is the counter-based
    one of csrc/philox.cuh (element (m, n), ld = N), kept values scaled by 1/(1-p), applied before the residual add."""
    from oracle.perceiver_ref import dropout_keep_mask
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    res = res if f32 else res.bfloat16()
    p, seed, offset, site = 0.1, 77, 16, 9
    keep = dropout_keep_mask(M, N, N, p, seed, offset, site).cuda().float() / (1.0 - p)
    d = ops.dropout_desc(p, seed, offset, site)
    out = ops.gemm(A, B, bias=bias, add=res, out_dtype=res.dtype, extra={"dropout": d})
    torch.cuda.synchronize()
    assert _rel(out, res.float() + keep * (ref + bias)) < (2e-3 if f32 else 1e-2)
    assert torch.equal(ops.dropout_mask(M, N, N, p, seed, offset, site).bool().cpu(),
                       dropout_keep_mask(M, N, N, p, seed, offset, site))
    if f32:
        return
    out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, extra={"dropout": d})
    torch.cuda.synchronize()
    assert _rel(out, keep * torch.nn.functional.gelu(ref + bias)) < 1e-2
    assert ((out == 0).float().mean() - p).abs() < 0.01
    pre = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    x = pre.float()
    dgelu = 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    out = ops.gemm(A, B, epilogue=ops.EPI_GELU_BWD, aux=pre, extra={"dropout": d})
    torch.cuda.synchronize()
    assert _rel(out, keep * ref * dgelu) < 1e-2


def _fold(W, b, gamma, beta):
    """Host-side folding of a LayerNorm into the Linear that consumes it (aggregator.Aggregator._pack.folded)."""
    w16 = (W * gamma[None, :]).bfloat16().contiguous()
    return w16, w16.float().sum(dim=1).contiguous(), (W @ beta + b).contiguous()


@pytest.mark.parametrize("M,N,K,mode", [(1374, 4096, 1024, "gelu"), (300, 512, 256, "linear"), (10992, 1024, 1024, "linear"),
                                        (777, 3072, 1024, "qkv")])
def test_gemm_layernorm_folded_consumer(M, N, K, mode):
    """LayerNorm folded into the consuming GEMM: A = raw bf16 rows, W' = W diag(gamma), per-row (sum, sum of squares)
    -> rstd (x W'^T - mean colsum(W')) + (W beta + b), against LayerNorm + Linear in fp32."""
    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g) * 3.0 + 0.7            # non-zero mean, non-unit variance
    x[:, 5] += 40.0                                                            # an outlier channel
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.3 * torch.randn(K, device="cuda", generator=g)
    beta = 0.2 * torch.randn(K, device="cuda", generator=g)
    eps = 1e-5
    xb, stats = ops.row_stats_cast(x)
    assert torch.equal(xb, x.bfloat16())
    assert torch.allclose(stats[:, 0], x.sum(1), rtol=1e-5, atol=1e-2) and torch.allclose(stats[:, 1], (x * x).sum(1), rtol=1e-5)
    w16, colsum, bias = _fold(W, b, gamma, beta)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (K,), gamma, beta, eps), W, b)
    extra = {"ln_stats": stats, "ln_colsum": colsum, "ln_eps": eps}
    if mode == "gelu":
        out = ops.gemm(xb, w16, bias=bias, epilogue=ops.EPI_GELU, extra=extra)
        ref = torch.nn.functional.gelu(ref)
    elif mode == "linear":
        out = ops.gemm(xb, w16, bias=bias, extra=extra)
    else:   # QKV epilogue without q/k norm + rope (DINOv2 stem form): bias and q scale only
        extra.update(tokens_per_frame=M, patch_start=5, grid_w=37, apply_qknorm_rope=0, q_scale=1.0)
        out = ops.gemm(xb, w16, bias=bias, epilogue=ops.EPI_QKV_ROPE, extra=extra)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1.5e-2


@pytest.mark.parametrize("M,N,K", [(1374, 1024, 4096), (10992, 1024, 1024), (130, 256, 128)])
def test_gemm_row_statistics_producer(M, N, K):
    """proj / fc2 epilogue of a LayerNorm-folded chain: besides C = add + gamma (acc + bias) (f32, in place) it leaves the
    bf16 copy of the new rows, their (sum, sum of squares) in `stats_out`, and clears `stats_zero`."""
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    expect = x + gamma * (ref + bias)
    stats = torch.zeros(M, 2, device="cuda")
    other = torch.full((M, 2), 7.0, device="cuda")
    copy = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(A, B, bias=bias, gamma=gamma, add=x, out=x, out_dtype=torch.float32,
                   extra={"stats_out": stats, "stats_zero": other, "c_copy_bf16": copy})
    torch.cuda.synchronize()
    assert _rel(out, expect) < 2e-3
    assert torch.equal(copy, out.bfloat16())
    assert torch.allclose(stats[:, 0], out.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[:, 1], (out * out).sum(1), rtol=1e-4)
    assert other.abs().max() == 0


@pytest.mark.parametrize("M,N,K,mode", [
    (10992, 1024, 4096, "lsres"),     # c2 fc2: 172 pair tiles on 74 pairs -> 24 tail tiles, 3 K slices each
    (10992, 4096, 1024, "gelu"),      # c2 fc1: 22 tail tiles, 3 slices of 5-6 k-blocks
    (1024, 16384, 4096, "gelu"),      # Perceiver fc1: 34 tail tiles, 2 slices
    (512, 4096, 2048, "linear"),      # fewer tiles than CTA pairs: every tile is a (split) tail tile
    (640, 512, 8192, "linear"),       # 6 tiles, 4 slices each
    (300, 256, 1024, "linear_f32"),   # single-CTA tiles (128 x 128), split tail
])
def test_gemm_k_sliced_tail_tiles(M, N, K, mode):
    """Tail tiles of the persistent schedule split along K across otherwise idle CTAs: same values as the unsplit
    schedule up to fp32 summation order, and bit-identical from run to run (partials are added in slice order)."""
    from vggt_qwen3_b200 import ops

    g, A, B, ref = _operands(M, N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    ops.gemm_tune(4)      # the split is off by default (it does not pay on B200); this test keeps the option honest

    def run():
        if mode == "lsres":
            gamma = torch.full((N,), 0.5, device="cuda")
            x = torch.ones(M, N, device="cuda")
            return ops.gemm(A, B, bias=bias, gamma=gamma, add=x, out=x, out_dtype=torch.float32), 1.0 + 0.5 * (ref + bias)
        if mode == "gelu":
            return ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU), torch.nn.functional.gelu(ref + bias)
        if mode == "linear_f32":
            return ops.gemm(A, B, bias=bias, out_dtype=torch.float32), ref + bias
        return ops.gemm(A, B, bias=bias), ref + bias

    out1, want = run()
    out1 = out1.clone()
    torch.cuda.synchronize()
    assert _rel(out1, want) < (2e-3 if out1.dtype == torch.float32 else 1e-2)
    try:
        for _ in range(3):
            out2, _ = run()
            torch.cuda.synchronize()
            assert torch.equal(out1, out2), "K-sliced tail is not run-to-run deterministic"
    finally:
        ops.gemm_tune(0)
