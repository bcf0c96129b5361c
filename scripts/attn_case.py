"""One attention launch configuration for ncu: python scripts/attn_case.py <n_seq> <L> <bounded 0|1> [emu]
Model-like inputs (q, k rows of norm 8 per head as after the QK LayerNorm, q scaled by log2(e)/8); 3 warm + 3 launches."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vggt_qwen3_b200 import ops  # noqa: E402

n_seq, L, bounded = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
emu = int(sys.argv[4]) if len(sys.argv) > 4 else -1
heads = 16
HD = heads * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
for blk, scale in ((0, math.log2(math.e) / 8.0), (1, 1.0)):
    v = qkv[:, blk * HD:(blk + 1) * HD].view(-1, heads, 64)
    v.copy_(v / v.norm(dim=-1, keepdim=True) * 8.0 * scale)
qkv = qkv.bfloat16()
out = torch.empty(n_seq * L, HD, device="cuda", dtype=torch.bfloat16)
if bounded:
    ops.attention_tune_bounded(1, emu)
for _ in range(6):
    if bounded:
        ops.attention_fwd_bounded(qkv, n_seq, L, heads, 12.0, out=out)
    else:
        ops.attention_fwd(qkv, n_seq, L, heads, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
