"""One-tile Perceiver attention at perceiver_small / c2 size for ncu and for CUDA-event timing against the multi-launch
formulation it replaces:  python scripts/pattn_case.py [p_dropout]   (B = 8, heads 8, 128 latents, 128 tokens, E = 4096)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vggt_qwen3_b200 import ops  # noqa: E402

p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
B, H, Lq, T, E = 8, 8, 128, 128, 4096
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B * Lq, E, device="cuda", generator=g).bfloat16()
kv = torch.randn(B * T, 2 * E, device="cuda", generator=g).bfloat16()
datt = torch.randn(B * Lq, E, device="cuda", generator=g).bfloat16()
drop = ops.dropout_desc(p, 1, 0, 0) if p > 0 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=20):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n)]
    for i in range(n):
        flush.zero_()                      # inputs leave L2 between iterations
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(n))[n // 2]


for _ in range(3):
    att, probs, _ = ops.perceiver_attention_fwd(q, kv, B, H, drop)
    dq, dkv = ops.perceiver_attention_bwd(q, kv, probs, datt, B, H, drop)
torch.cuda.synchronize()
t_f = timed(lambda: ops.perceiver_attention_fwd(q, kv, B, H, drop))
t_b = timed(lambda: ops.perceiver_attention_bwd(q, kv, probs, datt, B, H, drop))
print(f"ok fwd {t_f:.1f} us  bwd {t_b:.1f} us  (median of 20, L2 flushed, p = {p})", float(att.float().abs().mean()))
