"""A/B timing of the attention kernels in isolation (CUDA events, warm): checked wide kernel vs the bounded variant on
the frame (S x 1374) and global (S * 1374) shapes of c2 / c3 / c4, model-like scores.  usage: python scripts/attn_ab.py"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vggt_qwen3_b200 import ops  # noqa: E402


def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    heads = 16
    HD = heads * 64
    for n_seq, L in ((8, 1374), (32, 1374), (4, 8 * 1374), (1, 32 * 1374)):
        g = torch.Generator(device="cuda").manual_seed(0)
        qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
        for blk, scale in ((0, math.log2(math.e) / 8.0), (1, 1.0)):
            v = qkv[:, blk * HD:(blk + 1) * HD].view(-1, heads, 64)
            v.copy_(v / v.norm(dim=-1, keepdim=True) * 8.0 * scale)
        qkv = qkv.bfloat16()
        out = torch.empty(n_seq * L, HD, device="cuda", dtype=torch.bfloat16)
        flops = 4.0 * n_seq * heads * L * L * 64
        row = [f"{n_seq}x{L}"]
        ops.attention_tune(0, 4)
        ms = bench(lambda: ops.attention_fwd(qkv, n_seq, L, heads, out=out))
        row.append(f"checked {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s")
        for emu in (2, 3, 4):
            ops.attention_tune_bounded(1, emu)
            ms = bench(lambda: ops.attention_fwd_bounded(qkv, n_seq, L, heads, 12.0, out=out))
            row.append(f"bounded/emu{emu} {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s")
        print(" | ".join(row), flush=True)


if __name__ == "__main__":
    main()
