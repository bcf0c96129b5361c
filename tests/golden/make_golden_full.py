"""Full-size golden of the reference Perceiver (this container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_full.py

Runs the REFERENCE's own `PerceiverProjector` (src/models/projector_perceiver.py:53-82) at the size BASELINE.json names
(perceiver_small: 4096 / 128 latents / 8 heads / 6 layers / 16384, in 2048, out 2560; 1 227 684 352 parameters) on the
c2 shape (B = 8 samples x 128 context tokens), eval mode, fp32 on the CPU: forward + autograd backward.  The 4.9 GB of
gradients cannot be committed, so the fixture keeps, per tensor: its L2 norm, its dot products with four seeded +-1
probe vectors, and a strided sample of up to 4096 elements (`probe()` below; tests/test_perceiver_gpu.py re-derives the
same quantities from the CUDA result).  Weights and inputs are NOT stored: they are a pure function of the seeds
(`torch.manual_seed(0)` + the module constructor, which vggt_qwen3_b200.projector_perceiver mirrors draw for draw).
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

N_PROBES, N_SAMPLE = 4, 4096


def probe(t: torch.Tensor, index: int) -> np.ndarray:
    """[norm, dot(t, r_0..3), strided sample ...] with r_k = +-1 from a generator seeded by (index, k)."""
    flat = t.detach().reshape(-1).double().cpu()
    out = [float(flat.norm())]
    for k in range(N_PROBES):
        g = torch.Generator().manual_seed(1000 * index + k)
        r = torch.randint(0, 2, (flat.numel(),), generator=g, dtype=torch.int8).double() * 2 - 1
        out.append(float(flat @ r))
    stride = max(1, flat.numel() // N_SAMPLE)
    return np.concatenate([np.array(out), flat[::stride][:N_SAMPLE].numpy()])


def inputs(B: int = 8, T: int = 128, in_dim: int = 2048, Lq: int = 128, out_dim: int = 2560):
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(B, T, in_dim, generator=g)
    grad_out = torch.randn(B, Lq, out_dim, generator=g)
    return tokens, grad_out


def main():
    from src.models.projector_perceiver import PerceiverConfig, PerceiverProjector

    torch.set_num_threads(torch.get_num_threads())
    t0 = time.time()
    torch.manual_seed(0)
    m = PerceiverProjector(PerceiverConfig(), in_dim=2048, out_dim=2560).eval()
    tokens, grad_out = inputs()
    tokens.requires_grad_(True)
    out = m(tokens)
    out.backward(grad_out)
    blob = {"out": probe(out, 0), "grad_tokens": probe(tokens.grad, 1),
            "out_rows": out.detach()[:, ::16, :].reshape(-1, out.shape[-1])[:, ::4].numpy().astype(np.float32)}
    for i, (n, p) in enumerate(m.named_parameters()):
        blob["grad/" + n] = probe(p.grad, 2 + i)
    np.savez_compressed(HERE / "perceiver_ref_full.npz", **blob)
    print("perceiver_ref_full.npz", len(blob), "entries,", round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    main()
