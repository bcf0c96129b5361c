"""A/B of vq_gemm against cuBLAS (torch.matmul / F.linear, bf16) on the aggregator's GEMM shapes at the c2 and c3 row
counts (M = 10 992 and 43 968) and on the Perceiver's largest shapes; CUDA events, warm, L2 flushed between iterations
by cycling through operand copies larger than L2.  The library kernels carry their fused epilogues (bias only here; the
fused LayerScale/residual, GELU and QKV epilogues have no single cuBLAS counterpart).  usage: python scripts/gemm_ab.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vggt_qwen3_b200 import ops  # noqa: E402


def bench(fn, n_variants, iters=24):
    for i in range(4):
        fn(i % n_variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_variants)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    shapes = [("qkv", 3072, 1024), ("proj", 1024, 1024), ("fc1", 4096, 1024), ("fc2", 1024, 4096)]
    print("# shape M N K | vq_gemm us TF/s | cuBLAS us TF/s | ratio (vq / cuBLAS time)")
    for M in (10992, 43968):
        for name, N, K in shapes:
            nv = max(2, int(160e6 // (M * K * 2 + N * K * 2 + M * N * 2)) + 1)   # > 126 MB of L2 between reuses
            g = torch.Generator(device="cuda").manual_seed(0)
            As = [torch.randn(M, K, device="cuda", generator=g).bfloat16() for _ in range(nv)]
            Bs = [(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16() for _ in range(nv)]
            bias = torch.randn(N, device="cuda", generator=g)
            bias16 = bias.bfloat16()
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            t_vq = bench(lambda i: ops.gemm(As[i], Bs[i], bias=bias, out=out), nv)
            t_cb = bench(lambda i: torch.nn.functional.linear(As[i], Bs[i], bias16), nv)
            fl = 2.0 * M * N * K
            print(f"{name:5s} {M:6d} {N:5d} {K:5d} | {t_vq * 1e3:8.1f} {fl / t_vq / 1e9:7.1f} | {t_cb * 1e3:8.1f} "
                  f"{fl / t_cb / 1e9:7.1f} | {t_vq / t_cb:5.2f}", flush=True)
    for name, M, N, K in (("p_fc1", 1024, 16384, 4096), ("p_fc2", 1024, 4096, 16384), ("p_qkv", 1024, 8192, 4096)):
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        Bs = [(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16() for _ in range(2)]
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        t_vq = bench(lambda i: ops.gemm(A, Bs[i], out=out), 2)
        t_cb = bench(lambda i: torch.nn.functional.linear(A, Bs[i]), 2)
        fl = 2.0 * M * N * K
        print(f"{name:5s} {M:6d} {N:5d} {K:5d} | {t_vq * 1e3:8.1f} {fl / t_vq / 1e9:7.1f} | {t_cb * 1e3:8.1f} "
              f"{fl / t_cb / 1e9:7.1f} | {t_vq / t_cb:5.2f}", flush=True)


if __name__ == "__main__":
    main()
