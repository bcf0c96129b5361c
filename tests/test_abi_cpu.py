"""CPU checks of the C-ABI boundary: the library builds/loads without a GPU and exports every declared symbol."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "vq_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", text)) - {"vq_layer_done_fn"})


def test_library_exports_every_declared_symbol():
    from vggt_qwen3_b200.csrc.build import build

    lib = ctypes.CDLL(str(build()))
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vq_b200.h but not exported"


def test_python_binding_declares_every_symbol():
    from vggt_qwen3_b200 import _lib

    handle = _lib.lib()
    for n in _declared_symbols():
        assert getattr(handle, n).restype is not None or n == "vq_last_error"
    assert handle.vq_version() >= 1


def test_argument_validation_without_gpu():
    """Shape/alignment errors are reported through status codes + vq_last_error, never exceptions or crashes."""
    from vggt_qwen3_b200 import _lib

    handle = _lib.lib()
    d = _lib.GemmDesc()
    assert handle.vq_gemm(ctypes.byref(d), None) == -1
    assert b"M, N, K" in handle.vq_last_error()
    cfg = _lib.AggregatorCfg()
    assert handle.vq_aggregator_workspace_bytes(ctypes.byref(cfg)) == 0
    cfg.B, cfg.S, cfg.H, cfg.W, cfg.patch = 4, 8, 518, 518, 14
    cfg.D, cfg.heads, cfg.hidden, cfg.n_special, cfg.Kp = 1024, 16, 4096, 5, 592
    nbytes = handle.vq_aggregator_workspace_bytes(ctypes.byref(cfg))
    M = 4 * 8 * 1374
    assert nbytes >= M * (1024 * 4 + 1024 * 2 + 3072 * 2 + 1024 * 2 + 4096 * 2)
    assert nbytes < 2 * 1024 ** 3
    pc = _lib.PerceiverCfg()
    pc.B, pc.T, pc.in_dim, pc.E, pc.heads, pc.Lq, pc.layers, pc.ffn, pc.out_dim = 8, 128, 2048, 4096, 8, 128, 6, 16384, 2560
    pc.ln_eps = 1e-5
    assert handle.vq_perceiver_saved_bytes(ctypes.byref(pc)) > 0
    pc.heads = 7
    assert handle.vq_perceiver_saved_bytes(ctypes.byref(pc)) == 0


def test_product_path_refuses_cpu_tensors():
    from vggt_qwen3_b200.aggregator import Aggregator
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    agg = Aggregator(img_size=28, embed_dim=128, num_heads=2, depth=1, stem_depth=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        agg(torch.rand(1, 1, 3, 28, 28))
    proj = PerceiverProjector(PerceiverConfig(128, 64, 2, 1, 256, 0.0), 64, 64).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        proj(torch.randn(1, 8, 64))


def test_product_code_never_imports_the_oracle():
    for path in (ROOT / "vggt_qwen3_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_new_entry_points_validate_arguments_without_gpu():
    """f1 / f3 entry points and the attention variant switch: bad arguments come back as status codes + message."""
    from vggt_qwen3_b200 import _lib

    h = _lib.lib()
    assert h.vq_attention_tune(-1, -1) == 0
    assert h.vq_adamw_step(None, None, None, None, 0, None, 8, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert b"null pointer" in h.vq_last_error()
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert h.vq_adamw_step(p, p, p, p, 0, None, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert b"must be positive" in h.vq_last_error()
    assert h.vq_adamw_step(p, p, p, p, 0, None, 8, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0, None) == -1
    assert b"grad_scale" in h.vq_last_error()          # a zero scale is an error, not silently 1
    srcs = (ctypes.c_void_p * 2)(p, p)
    assert h.vq_adamw_reduce_step(p, p, p, srcs, 17, 0, None, 0, 8, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert b"gradient sources" in h.vq_last_error()
    assert h.vq_geom_head_forward(None, None, None, None, None, None, None, None, None, None, 1, 1, 37, 64, None) == -1
    assert h.vq_splice_features(None, 0, None, None, 0, None, 4, 1, 8, 64, 7, None, None) == -1
    d = _lib.DropoutDesc(); d.p = 1.5
    keep = (ctypes.c_uint8 * 64)()
    assert h.vq_dropout_mask(ctypes.cast(keep, ctypes.c_void_p), 8, 8, 8, ctypes.byref(d), None) == -1
    assert b"dropout p" in h.vq_last_error()
    assert h.vq_adamw_step(p, p, p, p, 0, None, 8, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert b"step counts from 1" in h.vq_last_error()
    assert h.vq_splice_insert(None, 0, None, 8, None, None, None, None, 1, 8, 64, 4, None) == -1
    assert h.vq_find_first_token(None, 0, 0, None, None) == -1
    # one-tile Perceiver attention: geometry is checked before anything touches the device
    assert h.vq_perceiver_attention_supported(128, 128, 4096, 8) == 1
    assert h.vq_perceiver_attention_supported(192, 128, 4096, 8) == 0
    assert h.vq_perceiver_attention_fwd(None, None, None, None, None, 1, 8, 128, 128, 4096, None, None) == -1
    assert b"null pointer" in h.vq_last_error()
    assert h.vq_perceiver_attention_fwd(p, p, p, None, p, 1, 8, 128, 256, 4096, None, None) == -1
    assert b"one-tile attention" in h.vq_last_error()
    assert h.vq_perceiver_attention_bwd(p, p, p, p, p, p, 1, 8, 128, 128, 4096 + 8, None, None) == -1
    assert b"head_dim" in h.vq_last_error()


def test_host_wrappers_refuse_cpu_tensors():
    """There is no CPU fallback: the Python layer raises instead of computing on the host."""
    from vggt_qwen3_b200 import ops
    from vggt_qwen3_b200.optim import FusedAdamW

    emb, mask = torch.zeros(1, 4, 8), torch.ones(1, 4, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.splice_insert(emb, mask, torch.zeros(1, 4, dtype=torch.int64), torch.zeros(1, 2, 8), 7)
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdamW([p], lr=1e-3).step()
    q, kv = torch.zeros(128, 512, dtype=torch.bfloat16), torch.zeros(128, 1024, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.perceiver_attention_fwd(q, kv, 1, 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.perceiver_attention_bwd(q, kv, torch.zeros(512, 128, dtype=torch.bfloat16), q, 1, 4)
