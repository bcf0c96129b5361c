"""GPU bring-up checks for the raw kernels (run on the B200 box via gpurun).

Each case runs in its own process (a device trap poisons the CUDA context):
    python scripts/gpu_bringup.py            # runs every case, prints one line per case
    python scripts/gpu_bringup.py <case>     # runs one case in-process
Compares against torch on the same GPU (fp32 math on the bf16-rounded inputs).
"""
from __future__ import annotations

import math
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _rel(a, b):
    import torch

    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _cos(a, b):
    import torch

    a = a.float().flatten()
    b = b.float().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def case_gemm(M, N, K, a_major=0, b_major=0, epi="linear", out_f32=False, timing=False):
    import torch

    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    ref = A.float() @ B.float().t()
    a_in = A.t().contiguous() if a_major else A
    b_in = B.t().contiguous() if b_major else B
    kw = {}
    bias = torch.randn(N, device="cuda", generator=g)
    if epi == "linear":
        kw = dict(bias=bias)
        ref = ref + bias
    elif epi == "gelu":
        aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(bias=bias, epilogue=ops.EPI_GELU, aux=aux)
        pre = ref + bias
        ref = torch.nn.functional.gelu(pre)
    elif epi == "gelu_noaux":
        kw = dict(bias=bias, epilogue=ops.EPI_GELU)
        ref = torch.nn.functional.gelu(ref + bias)
    elif epi == "ls_bf16":      # LayerScale-d branch output, bf16, no residual (deferred-residual schedule)
        gamma = torch.rand(N, device="cuda", generator=g)
        kw = dict(bias=bias, gamma=gamma)
        ref = gamma * (ref + bias)
    elif epi == "lsres":
        gamma = torch.rand(N, device="cuda", generator=g)
        res = torch.randn(M, N, device="cuda", generator=g)
        if not out_f32:
            res = res.bfloat16()
        kw = dict(bias=bias, gamma=gamma, add=res)
        ref = res.float() + gamma * (ref + bias)
    elif epi == "gelu_bwd":
        pre = torch.randn(M, N, device="cuda", generator=g).bfloat16()
        kw = dict(epilogue=ops.EPI_GELU_BWD, aux=pre)
        x = pre.float()
        cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
        pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
        ref = ref * (cdf + x * pdf)
    kw["out_dtype"] = torch.float32 if out_f32 else torch.bfloat16
    out = ops.gemm(a_in, b_in, a_major=a_major, b_major=b_major, **kw)
    torch.cuda.synchronize()
    rel = _rel(out, ref)
    msg = f"rel={rel:.3e} cos={_cos(out, ref):.6f}"
    if epi == "gelu":
        msg += f" aux_rel={_rel(aux, pre):.3e}"
    ok = rel < (2e-3 if out_f32 else 1e-2)
    if timing:
        for _ in range(3):
            ops.gemm(a_in, b_in, a_major=a_major, b_major=b_major, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.gemm(a_in, b_in, a_major=a_major, b_major=b_major, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        msg += f" time={ms*1e3:.1f}us {2*M*N*K/ms/1e9:.1f}TF/s"
        e0.record()
        for _ in range(n):
            torch.nn.functional.linear(A, B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        msg += f" | cublas {ms*1e3:.1f}us {2*M*N*K/ms/1e9:.1f}TF/s"
    return ok, msg


def case_qkv_rope(frames, grid, heads, timing=False):
    """QKV GEMM with the fused per-head QK LayerNorm + 2-D RoPE + q pre-scale epilogue against torch fp32."""
    import torch

    from vggt_qwen3_b200 import ops

    def rope_tables(max_pos, head_dim, freq=100.0):   # [UPSTREAM] vggt/layers/rope.py: inv_freq_j = freq^(-2j/(hd/2))
        half = head_dim // 2
        inv = 1.0 / (freq ** (torch.arange(0, half, 2, dtype=torch.float32) / half))
        ang = torch.arange(max_pos + 1, dtype=torch.float32)[:, None] * inv[None, :]
        return ang.cos(), ang.sin()

    def rope_1d(t, pos, cos, sin):                    # out = t * cos + rotate_half(t) * sin, cos = cat(theta, theta)
        c, s_ = torch.cat([cos[pos], cos[pos]], -1), torch.cat([sin[pos], sin[pos]], -1)
        h = t.shape[-1] // 2
        return t * c + torch.cat([-t[..., h:], t[..., :h]], -1) * s_

    def rope_2d(t, pos, cos, sin):                    # first half of the head rotates with y, second half with x
        half = t.shape[-1] // 2
        py, px = pos[..., 0][:, None, :], pos[..., 1][:, None, :]
        return torch.cat([rope_1d(t[..., :half], py, cos, sin), rope_1d(t[..., half:], px, cos, sin)], -1)

    g = torch.Generator(device="cuda").manual_seed(5)
    D, P = heads * 64, 5 + grid * grid
    M = frames * P
    x = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    bias = torch.randn(3 * D, device="cuda", generator=g) * 0.1
    qn_w, kn_w = (torch.rand(64, device="cuda", generator=g) + 0.5 for _ in range(2))
    qn_b, kn_b = (torch.randn(64, device="cuda", generator=g) * 0.1 for _ in range(2))
    cos, sin = (t.cuda().contiguous() for t in rope_tables(grid, 64))
    q_scale = 1.4426950408889634 / 8.0
    extra = dict(tokens_per_frame=P, patch_start=5, grid_w=grid, apply_qknorm_rope=1, qn_w=qn_w, qn_b=qn_b, kn_w=kn_w,
                 kn_b=kn_b, qk_eps=1e-5, q_scale=q_scale, rope_cos=cos, rope_sin=sin, rope_positions=grid + 1)
    out = ops.gemm(x, w, bias=bias, epilogue=ops.EPI_QKV_ROPE, extra=extra)
    torch.cuda.synchronize()
    qkv = (x.float() @ w.float().t() + bias).view(frames, P, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q = torch.nn.functional.layer_norm(qkv[0], (64,), qn_w, qn_b, 1e-5)
    k = torch.nn.functional.layer_norm(qkv[1], (64,), kn_w, kn_b, 1e-5)
    pos = torch.zeros(P, 2, dtype=torch.long, device="cuda")
    yy, xx = torch.meshgrid(torch.arange(grid, device="cuda"), torch.arange(grid, device="cuda"), indexing="ij")
    pos[5:, 0], pos[5:, 1] = yy.flatten() + 1, xx.flatten() + 1
    pos = pos[None].expand(frames, P, 2)
    q = rope_2d(q, pos, cos, sin) * q_scale
    k = rope_2d(k, pos, cos, sin)
    ref = torch.stack([q, k, qkv[2]], 0).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
    rel = _rel(out, ref)
    msg = f"rel={rel:.3e} cos={_cos(out, ref):.6f}"
    if timing:
        for _ in range(3):
            ops.gemm(x, w, bias=bias, epilogue=ops.EPI_QKV_ROPE, extra=extra, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.gemm(x, w, bias=bias, epilogue=ops.EPI_QKV_ROPE, extra=extra, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        msg += f" time={ms*1e3:.1f}us {2*M*3*D*D/ms/1e9:.1f}TF/s"
    return rel < 1e-2, msg


def case_attn(n_seq, L, heads, timing=False, qscale=2.0):
    import torch

    from vggt_qwen3_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    HD = heads * 64
    qkv = torch.randn(n_seq * L, 3 * HD, device="cuda", generator=g)
    qkv[:, :HD] *= qscale  # sharper softmax
    qkv = qkv.bfloat16()
    q, k, v = (qkv[:, i * HD:(i + 1) * HD].float().view(n_seq, L, heads, 64).transpose(1, 2) for i in range(3))
    # kernel convention: q is pre-scaled by log2(e)/sqrt(d); softmax base 2  => plain softmax of q.k * ln2
    s = (q @ k.transpose(-1, -2)) * math.log(2.0)
    ref = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(n_seq * L, HD)
    out = ops.attention_fwd(qkv, n_seq, L, heads)
    torch.cuda.synchronize()
    rel = _rel(out, ref)
    msg = f"rel={rel:.3e} cos={_cos(out, ref):.6f}"
    ok = rel < 2e-2
    if timing:
        for _ in range(3):
            ops.attention_fwd(qkv, n_seq, L, heads)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            ops.attention_fwd(qkv, n_seq, L, heads)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * n_seq * heads * L * L * 64
        msg += f" time={ms*1e3:.1f}us {fl/ms/1e9:.1f}TF/s"
        qq, kk, vv = (qkv[:, i * HD:(i + 1) * HD].view(n_seq, L, heads, 64).transpose(1, 2) for i in range(3))
        for _ in range(2):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e0.record()
        for _ in range(n):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        msg += f" | sdpa {ms*1e3:.1f}us {fl/ms/1e9:.1f}TF/s"
    return ok, msg


CASES = {
    # --- GEMM: K-major x K-major, shapes of increasing raggedness
    "gemm_small": lambda: case_gemm(128, 128, 64),
    "gemm_k256": lambda: case_gemm(128, 256, 256),
    "gemm_ragged": lambda: case_gemm(1374, 1024, 1024),
    "gemm_kragged": lambda: case_gemm(300, 1024, 592),
    "gemm_f32": lambda: case_gemm(777, 512, 320, out_f32=True),
    "gemm_gelu": lambda: case_gemm(1374, 4096, 1024, epi="gelu"),
    "gemm_lsres_f32": lambda: case_gemm(1374, 1024, 4096, epi="lsres", out_f32=True),
    "gemm_lsres_bf16": lambda: case_gemm(1024, 4096, 4096, epi="lsres"),
    "gemm_gelu_bwd": lambda: case_gemm(1024, 2048, 512, epi="gelu_bwd"),
    # --- operand majors
    "gemm_bmn": lambda: case_gemm(1024, 1024, 512, b_major=1),
    "gemm_amn": lambda: case_gemm(1024, 1024, 512, a_major=1),
    "gemm_abmn": lambda: case_gemm(4096, 2048, 1024, a_major=1, b_major=1, out_f32=True),
    # --- timing at the aggregator shapes (c2: M = 8*1374)
    "gemm_t_qkvshape": lambda: case_gemm(10992, 3072, 1024, timing=True),
    "gemm_t_fc1": lambda: case_gemm(10992, 4096, 1024, epi="gelu", timing=True),
    "gemm_t_fc1_noaux": lambda: case_gemm(10992, 4096, 1024, epi="gelu_noaux", timing=True),   # as the aggregator runs it
    "gemm_t_proj_bf16": lambda: case_gemm(10992, 1024, 1024, epi="ls_bf16", timing=True),      # deferred-residual schedule
    "gemm_t_fc2_bf16": lambda: case_gemm(10992, 1024, 4096, epi="ls_bf16", timing=True),
    "gemm_t_fc2": lambda: case_gemm(10992, 1024, 4096, epi="lsres", out_f32=True, timing=True),
    "gemm_t_proj": lambda: case_gemm(10992, 1024, 1024, epi="lsres", out_f32=True, timing=True),
    "gemm_t_8k": lambda: case_gemm(8192, 8192, 8192, timing=True),
    # --- c3 (M = 32*1374)
    "gemm_t3_qkv": lambda: case_gemm(43968, 3072, 1024, timing=True),
    "gemm_t3_fc1": lambda: case_gemm(43968, 4096, 1024, epi="gelu", timing=True),
    "gemm_t3_fc2": lambda: case_gemm(43968, 1024, 4096, epi="lsres", out_f32=True, timing=True),
    "gemm_t3_proj": lambda: case_gemm(43968, 1024, 1024, epi="lsres", out_f32=True, timing=True),
    "gemm_qkvrope_small": lambda: case_qkv_rope(3, 6, 2),
    "gemm_t_qkvrope": lambda: case_qkv_rope(8, 37, 16, timing=True),
    "gemm_t3_qkvrope": lambda: case_qkv_rope(32, 37, 16, timing=True),
    # --- attention
    "attn_128": lambda: case_attn(1, 128, 1),
    "attn_256": lambda: case_attn(1, 256, 2),
    "attn_ragged_small": lambda: case_attn(2, 200, 2),
    "attn_l50": lambda: case_attn(3, 50, 1),
    "attn_l64": lambda: case_attn(2, 64, 2),
    "attn_l65": lambda: case_attn(2, 65, 2),
    "attn_l129": lambda: case_attn(2, 129, 1),
    "attn_l300": lambda: case_attn(2, 300, 3),
    "attn_l1374_sharp": lambda: case_attn(2, 1374, 2, qscale=12.0),
    "attn_l2000_flat": lambda: case_attn(1, 2000, 2, qscale=0.05),
    "attn_frame": lambda: case_attn(8, 1374, 16, timing=True),
    "attn_frame32": lambda: case_attn(32, 1374, 16, timing=True),
    "attn_global8": lambda: case_attn(1, 8 * 1374, 16, timing=True),
    "attn_global8x4": lambda: case_attn(4, 8 * 1374, 16, timing=True),
    # model-like score magnitudes (LayerNorm-ed q, k: |s| of a few units)
    "attn_frame32_m": lambda: case_attn(32, 1374, 16, timing=True, qscale=0.25),
    "attn_global8x4_m": lambda: case_attn(4, 8 * 1374, 16, timing=True, qscale=0.25),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        name = sys.argv[1]
        ok, msg = CASES[name]()
        print(f"{'PASS' if ok else 'FAIL'} {name}: {msg}", flush=True)
        sys.exit(0 if ok else 1)
    names = [n for n in CASES if not sys.argv[1:] or any(n.startswith(p) for p in sys.argv[1:])]
    n_fail = 0
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=75)
            out = (p.stdout.strip().splitlines() or ["<no output>"])[-1]
            if p.returncode != 0:
                n_fail += 1
                if not out.startswith("FAIL"):
                    out = f"ERROR {name}: rc={p.returncode} " + (p.stdout + p.stderr).strip()[-600:].replace("\n", " | ")
        except subprocess.TimeoutExpired:
            n_fail += 1
            out = f"TIMEOUT {name}"
        print(f"{out}  [{time.time()-t0:.1f}s]", flush=True)
        if out.startswith("TIMEOUT"):
            break  # a hung kernel: do not burn GPU time on the remaining cases
    print(f"bringup: {len(names)-n_fail}/{len(names)} passed", flush=True)
    sys.exit(1 if n_fail else 0)


if __name__ == "__main__":
    main()
