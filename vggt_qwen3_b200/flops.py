"""Algorithmic FLOP model of the path: 2*M*N*K per GEMM, 4*Nq*Nkv*D per attention,
no recompute, no credit for pruning.  Pure arithmetic (no tensors); bench.py's roofline figures are computed from it."""
from __future__ import annotations


def aggregator_flops_per_frame(S: int, H: int = 518, W: int = 518, *, patch: int = 14, D: int = 1024, heads: int = 16,
                               mlp_ratio: float = 4.0, depth: int = 24, stem_depth: int = 24, registers: int = 4) -> float:
    """VGGT Aggregator forward, per frame of an S-frame scene: patch-embed conv + (stem_depth + 2*depth) blocks of
    linear layers + stem/frame attention over P tokens + global attention over S*P tokens.
    3048.0 GF (S=1), 4347.1 GF (S=8), 8801.2 GF (S=32) at 518 px."""
    gh, gw = H // patch, W // patch
    P = 1 + registers + gh * gw
    hidden = int(D * mlp_ratio)
    lin = 2.0 * P * (D * 3 * D + D * D + 2 * D * hidden)
    att = 4.0 * P * P * D
    conv = 2.0 * gh * gw * (3 * patch * patch) * D
    return conv + (stem_depth + 2 * depth) * lin + (stem_depth + depth) * att + depth * att * S


def perceiver_flops_per_sample(T: int = 128, *, in_dim: int = 2048, E: int = 4096, Lq: int = 128, heads: int = 8,
                               layers: int = 6, ffn: int = 16384, out_dim: int = 2560) -> float:
    """PerceiverProjector forward per sample (315.68 GF for perceiver_small at T = 128); backward = 2x forward."""
    hd = E // heads
    per_layer = 2.0 * Lq * E * E + 2.0 * T * E * 2 * E + 2.0 * Lq * E * E + 4.0 * Lq * T * hd * heads + 4.0 * Lq * E * ffn
    return 2.0 * T * in_dim * E + layers * per_layer + 2.0 * Lq * E * out_dim


def workload_flops(B: int, S: int, backward: bool, num_vis_tokens: int = 128) -> dict:
    """Per step per GPU: B scenes of S frames through the aggregator, B samples through the Perceiver (x3 with backward)."""
    agg = aggregator_flops_per_frame(S) * B * S
    perc = perceiver_flops_per_sample(num_vis_tokens) * B * (3.0 if backward else 1.0)
    return dict(aggregator=agg, perceiver=perc, total=agg + perc)
