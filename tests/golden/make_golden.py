"""Generate the committed golden fixtures by running the REFERENCE's own code (this container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Needs /root/reference (read-only, absent on the GPU box) and transformers.  Writes small .npz files next to
this script; tests read only the .npz files.

  perceiver_ref_small.npz   reference `PerceiverProjector` (src/models/projector_perceiver.py:53-82), eval mode:
                            its own seed-0 init (biases / LN affine / latents perturbed so those paths are live),
                            tokens, output, grad_out, every parameter gradient and the token gradient.
  perceiver_ref_train.npz   the same reference class in TRAIN mode (dropout 0.1, src/train/train_sft.py:196) with
                            torch.nn.functional.dropout patched to apply the counter-based masks of csrc/philox.cuh
                            in call order (per layer: attention probabilities, post-attention, MLP, post-MLP): pins
                            WHERE the reference drops and the site numbering; output + all gradients.
  encode_images_ref.npz     reference `VGGTQwen3VLM.encode_images` (src/models/vggt_qwen3_vlm.py:128-162) called
                            unbound on a stub whose aggregator returns an index-tagged list: pins the slice /
                            token-order contract (first num_vis_tokens rows of the view-flattened last entry).
  insert_splice_ref.npz     reference `_insert_vision_tokens` (src/inference/qa_inference.py:119-145).
  dinov2_hf_tiny.npz        HuggingFace Dinov2WithRegistersModel (tiny random config): weights, image, output —
                            pins the DINOv2-stem part of oracle/aggregator_ref.py.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF))
sys.path.insert(0, str(HERE.parent.parent))   # oracle/ (mask function of the train-mode golden)


def perceiver_golden():
    from src.models.projector_perceiver import PerceiverConfig, PerceiverProjector

    torch.manual_seed(0)
    cfg = PerceiverConfig(latent_dim=128, num_latents=64, num_heads=2, num_layers=2, ffn_dim=256, dropout=0.1)
    m = PerceiverProjector(cfg, in_dim=64, out_dim=96).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") or n.endswith("in_proj_bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif "norm" in n and n.endswith("weight"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
        m.latents.mul_(25.0)  # std 0.5: keeps the attention logits away from the uniform regime
    B, T = 3, 64
    tokens = torch.randn(B, T, 64, generator=g, requires_grad=True)
    out = m(tokens)
    grad_out = torch.randn(out.shape, generator=g)
    out.backward(grad_out)
    blob = {"cfg": np.array([cfg.latent_dim, cfg.num_latents, cfg.num_heads, cfg.num_layers, cfg.ffn_dim, 64, 96]),
            "tokens": tokens.detach().numpy(), "out": out.detach().numpy(), "grad_out": grad_out.numpy(),
            "grad_tokens": tokens.grad.numpy()}
    for n, p in m.named_parameters():
        blob["param/" + n] = p.detach().numpy()
        blob["grad/" + n] = p.grad.numpy()
    np.savez_compressed(HERE / "perceiver_ref_small.npz", **blob)
    print("perceiver_ref_small.npz", out.shape, float(out.detach().abs().mean()))


def perceiver_train_golden():
    """Reference PerceiverProjector under .train(): every dropout the reference executes (nn.Dropout modules and the
    attention-probability dropout inside nn.MultiheadAttention all go through torch.nn.functional.dropout) is replaced
    by the deterministic mask of its call index, so the result can be reproduced by the oracle and by the kernels."""
    import torch.nn.functional as F

    from oracle import perceiver_ref as P
    from src.models.projector_perceiver import PerceiverConfig, PerceiverProjector

    torch.manual_seed(0)
    cfg = PerceiverConfig(latent_dim=128, num_latents=64, num_heads=2, num_layers=2, ffn_dim=256, dropout=0.1)
    m = PerceiverProjector(cfg, in_dim=64, out_dim=96).train()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") or n.endswith("in_proj_bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif "norm" in n and n.endswith("weight"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
        m.latents.mul_(25.0)
    B, T = 3, 72   # T not a multiple of 64: the attention-probability mask is indexed with the padded row length
    seed, offset = 20240611, 12
    calls = []
    real_dropout = F.dropout

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        assert training and abs(p - 0.1) < 1e-9
        site = len(calls)
        calls.append(tuple(x.shape))
        rows = x.numel() // x.shape[-1]
        cols = x.shape[-1]
        ld = (cols + 63) // 64 * 64 if site % 4 == 0 else cols
        keep = P.dropout_keep_mask(rows, cols, ld, p, seed, offset, site).reshape(x.shape)
        return x * keep.to(x.dtype) * (1.0 / (1.0 - p))

    F.dropout = fake_dropout
    try:
        tokens = torch.randn(B, T, 64, generator=g, requires_grad=True)
        out = m(tokens)
        grad_out = torch.randn(out.shape, generator=g)
        out.backward(grad_out)
    finally:
        F.dropout = real_dropout
    # call order per layer: attention probabilities [B*H, Lq, T], post-attention [B, Lq, E], MLP [B, Lq, F], post-MLP
    assert calls == [(B * 2, 64, T), (B, 64, 128), (B, 64, 256), (B, 64, 128)] * 2, calls
    blob = {"cfg": np.array([cfg.latent_dim, cfg.num_latents, cfg.num_heads, cfg.num_layers, cfg.ffn_dim, 64, 96]),
            "dropout": np.array([0.1, seed, offset]), "tokens": tokens.detach().numpy(), "out": out.detach().numpy(),
            "grad_out": grad_out.numpy(), "grad_tokens": tokens.grad.numpy()}
    for n, p in m.named_parameters():
        blob["param/" + n] = p.detach().numpy()
        blob["grad/" + n] = p.grad.numpy()
    np.savez_compressed(HERE / "perceiver_ref_train.npz", **blob)
    print("perceiver_ref_train.npz", out.shape, float(out.detach().abs().mean()))


def encode_images_golden():
    from src.models.vggt_qwen3_vlm import VGGTQwen3VLM

    B, V, P, C = 2, 3, 21, 16
    num_vis = 30  # > P: the slice continues into frame 1 of the view-flattened sequence

    class Agg:
        def __call__(self, images):
            # value encodes (b, v, p): makes the token order checkable exactly
            b = torch.arange(B).view(B, 1, 1, 1) * 10000
            v = torch.arange(V).view(1, V, 1, 1) * 100
            p = torch.arange(P).view(1, 1, P, 1)
            last = (b + v + p).expand(B, V, P, C).to(torch.bfloat16)
            return [torch.zeros_like(last), last], 5

    class Proj(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.in_proj = torch.nn.Linear(C, 4)

        def forward(self, x):
            return x[..., :1].float()  # expose which rows were selected, in order

    stub = types.SimpleNamespace(vision_model=types.SimpleNamespace(aggregator=Agg()), num_vis_tokens=num_vis,
                                 projector=Proj())
    images = torch.zeros(B, V, 3, 14, 14)
    out = VGGTQwen3VLM.encode_images(stub, images)
    assert out.requires_grad is False
    np.savez_compressed(HERE / "encode_images_ref.npz", shape=np.array([B, V, P, C, num_vis]), selected=out.numpy())
    print("encode_images_ref.npz", out.shape)


def insert_splice_golden():
    # the module itself does not import under transformers 5.x (`load_sharded_checkpoint` was removed), so the
    # reference's function is compiled straight from its source file, unmodified
    import ast
    from typing import Tuple

    src = (REF / "src/inference/qa_inference.py").read_text()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "_insert_vision_tokens")
    ns = {"torch": torch, "Tuple": Tuple}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "qa_inference.py", "exec"), ns)
    _insert_vision_tokens = ns["_insert_vision_tokens"]

    g = torch.Generator().manual_seed(3)
    L, H, n, image_id = 12, 8, 5, 99
    ids = torch.randint(0, 50, (1, L), generator=g)
    ids[0, 4] = image_id
    ids[0, 9] = image_id  # only the first one is expanded
    mask = torch.ones(1, L, dtype=torch.long)
    mask[0, :2] = 0
    emb = torch.randn(1, L, H, generator=g)
    vis = torch.randn(1, n, H, generator=g)
    new_emb, new_mask = _insert_vision_tokens(ids, mask, emb, vis, image_id)
    ids2 = ids.clone()
    ids2[ids2 == image_id] = 1
    same_emb, same_mask = _insert_vision_tokens(ids2, mask, emb, vis, image_id)
    assert same_emb is emb and same_mask is mask
    np.savez_compressed(HERE / "insert_splice_ref.npz", ids=ids.numpy(), mask=mask.numpy(), emb=emb.numpy(),
                        vis=vis.numpy(), image_id=np.array(image_id), new_emb=new_emb.numpy(), new_mask=new_mask.numpy())
    print("insert_splice_ref.npz", new_emb.shape)


def dinov2_golden():
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    torch.manual_seed(0)
    cfg = Dinov2WithRegistersConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, mlp_ratio=4,
                                    image_size=56, patch_size=14, num_register_tokens=4, layerscale_value=1.0,
                                    hidden_act="gelu", layer_norm_eps=1e-6, qkv_bias=True)
    m = Dinov2WithRegistersModel(cfg).eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lambda1" in n:  # LayerScale
                p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.5)
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif "register_tokens" in n or "cls_token" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
    x = torch.randn(2, 3, 56, 56, generator=g)
    with torch.no_grad():
        y = m(pixel_values=x).last_hidden_state
    blob = {"x": x.numpy(), "y": y.numpy()}
    for n, p in m.state_dict().items():
        blob["hf/" + n] = p.numpy()
    np.savez_compressed(HERE / "dinov2_hf_tiny.npz", **blob)
    print("dinov2_hf_tiny.npz", y.shape)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":   # add the train-mode golden without regenerating the others
        perceiver_train_golden()
        sys.exit(0)
    perceiver_golden()
    perceiver_train_golden()
    encode_images_golden()
    insert_splice_golden()
    dinov2_golden()
