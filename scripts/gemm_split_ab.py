"""Interleaved A/B of the GEMM's K-sliced tail (same process, alternating settings, min of 5 rounds): fused epilogues as
the aggregator / Perceiver run them.  usage: python scripts/gemm_split_ab.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vggt_qwen3_b200 import ops  # noqa: E402


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    cases = [("fc2 c2 lsres", 10992, 1024, 4096, "lsres"), ("fc1 c2 gelu", 10992, 4096, 1024, "gelu"),
             ("proj c2 lsres", 10992, 1024, 1024, "lsres"), ("fc2 c3 lsres", 43968, 1024, 4096, "lsres"),
             ("fc1 c3 gelu", 43968, 4096, 1024, "gelu"), ("p_fc1 gelu", 1024, 16384, 4096, "gelu"),
             ("p_fc2 lin", 1024, 4096, 16384, "linear"), ("p_dgrad", 1024, 4096, 4096, "linear")]
    for name, M, N, K, mode in cases:
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        gamma = torch.rand(N, device="cuda", generator=g)
        x = torch.zeros(M, N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        if mode == "lsres":
            fn = lambda: ops.gemm(A, B, bias=bias, gamma=gamma, add=x, out=x, out_dtype=torch.float32)
        elif mode == "gelu":
            fn = lambda: ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, out=out)
        else:
            fn = lambda: ops.gemm(A, B, bias=bias, out=out)
        best = {}
        for _ in range(5):
            for s in (0, 2, 3, 4):
                ops.gemm_tune(s)
                best[s] = min(best.get(s, 1e9), timeit(fn))
        fl = 2.0 * M * N * K
        print(f"{name:14s} " + " | ".join(f"split<={s}: {t:7.1f} us {fl / t / 1e6:7.1f} TF/s" for s, t in best.items()), flush=True)
    ops.gemm_tune(4)


if __name__ == "__main__":
    main()
