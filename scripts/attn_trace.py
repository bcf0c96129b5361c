"""Debug: dump clock64() timestamps of one CTA of the attention kernel (MMA thread + one lane of each softmax WG)."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vggt_qwen3_b200 import _lib, ops
lib = _lib.lib()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10992
qkv = torch.randn(L, 3 * 1024, device="cuda").bfloat16()
trace = torch.zeros(3 * 32 * 8, dtype=torch.int64, device="cuda")
ops.attention_fwd(qkv, 1, L, 16)
lib.vq_debug_set_attention_trace.argtypes = [ctypes.c_void_p]
lib.vq_debug_set_attention_trace(ctypes.c_void_p(trace.data_ptr()))
ops.attention_fwd(qkv, 1, L, 16)
torch.cuda.synchronize()
lib.vq_debug_set_attention_trace(None)
t = trace.cpu().view(3, 32, 8)
t0 = int(t[0, 4, 0])
names = {0: ["iter", "kfull", "sfree0", "sfree1", "vfull", "pfull0", "pfull1", "-"],
         1: ["iter", "sfull", "sfree^", "premax", "ppsync", "expend", "odone", "pfull^"]}
for j in range(4, 12):
    print(f"j={j} MMA  " + " ".join(f"{names[0][e]}={int(t[0,j,e])-t0:6d}" for e in range(7)))
    for r in (1, 2):
        print(f"j={j} SM{r-1}  " + " ".join(f"{names[1][e]}={int(t[r,j,e])-t0:6d}" for e in range(8)))
