"""Epilogue cost of the aggregator GEMMs at the c2 / c3 row counts, one process, interleaved, min of 5 rounds:
plain bias epilogue vs the fused ones (GELU, LayerScale + f32 residual, QKV LayerNorm + RoPE).  usage: python scripts/gemm_epi_ab.py"""
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
    for M in (10992, 43968):
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda r, c, s=1.0: (torch.randn(r, c, device="cuda", generator=g) * s).bfloat16()
        A1, A4 = mk(M, 1024), mk(M, 4096)
        Wqkv, Wproj, Wfc1, Wfc2 = mk(3072, 1024, 0.03), mk(1024, 1024, 0.03), mk(4096, 1024, 0.03), mk(1024, 4096, 0.015)
        b3, b1, b4 = (torch.randn(n, device="cuda", generator=g) for n in (3072, 1024, 4096))
        gam = torch.rand(1024, device="cuda", generator=g)
        x = torch.zeros(M, 1024, device="cuda")
        o3, o1, o4 = (torch.empty(M, n, device="cuda", dtype=torch.bfloat16) for n in (3072, 1024, 4096))
        qn = [torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")]
        pos = torch.arange(38, device="cuda", dtype=torch.float32)[:, None] * torch.rand(1, 16, device="cuda", generator=g)
        cos, sin = pos.cos().contiguous(), pos.sin().contiguous()
        rope = dict(tokens_per_frame=1374, patch_start=5, grid_w=37, apply_qknorm_rope=1, qn_w=qn[0], qn_b=qn[1], kn_w=qn[0],
                    kn_b=qn[1], qk_eps=1e-5, q_scale=0.18, rope_cos=cos, rope_sin=sin, rope_positions=38)
        cases = {
            "qkv  plain": lambda: ops.gemm(A1, Wqkv, bias=b3, out=o3),
            "qkv  ln+rope": lambda: ops.gemm(A1, Wqkv, bias=b3, out=o3, epilogue=ops.EPI_QKV_ROPE, extra=rope),
            "proj plain bf16": lambda: ops.gemm(A1, Wproj, bias=b1, out=o1),
            "proj ls+res f32": lambda: ops.gemm(A1, Wproj, bias=b1, gamma=gam, add=x, out=x, out_dtype=torch.float32),
            "fc1  plain": lambda: ops.gemm(A1, Wfc1, bias=b4, out=o4),
            "fc1  gelu": lambda: ops.gemm(A1, Wfc1, bias=b4, out=o4, epilogue=ops.EPI_GELU),
            "fc2  plain bf16": lambda: ops.gemm(A4, Wfc2, bias=b1, out=o1),
            "fc2  ls+res f32": lambda: ops.gemm(A4, Wfc2, bias=b1, gamma=gam, add=x, out=x, out_dtype=torch.float32),
        }
        best = {}
        for _ in range(5):
            for k, fn in cases.items():
                best[k] = min(best.get(k, 1e9), timeit(fn))
        for k, t in best.items():
            print(f"M={M:6d} {k:16s} {t:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
