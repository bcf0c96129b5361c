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
    vggt/layers/attention.py q_norm / k_norm, vggt/layers/rope.py) against the oracle's rope_2d on torch fp32."""
    from scripts.gpu_bringup import case_qkv_rope

    ok, msg = case_qkv_rope(frames, grid, heads)
    assert ok, msg


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
