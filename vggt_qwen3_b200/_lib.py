"""ctypes binding of libvq_b200.so (C ABI declared in include/vq_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# VQ_B200_LIB: load an experimental build of the same library (kernel tuning); the default is the in-tree build
LIB_PATH = Path(os.environ["VQ_B200_LIB"]) if os.environ.get("VQ_B200_LIB") else _PKG / "libvq_b200.so"

VQ_BF16, VQ_F32 = 0, 1
EPI_LINEAR, EPI_GELU, EPI_GELU_BWD, EPI_QKV_ROPE, EPI_PATCH = 0, 1, 2, 3, 4

_f32p = C.c_void_p


class DropoutDesc(C.Structure):
    """Mirror of `vq_dropout_desc`."""

    _fields_ = [("p", C.c_float), ("site", C.c_uint32), ("seed", C.c_uint64), ("offset", C.c_uint64)]


class GemmDesc(C.Structure):
    """Mirror of `vq_gemm_desc` (include/vq_b200.h)."""

    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("c_dtype", C.c_int32),
        ("Z0", C.c_int32), ("Z1", C.c_int32),
        ("a_r0", C.c_int64), ("a_c0", C.c_int64), ("a_r1", C.c_int64), ("a_c1", C.c_int64),
        ("b_r0", C.c_int64), ("b_c0", C.c_int64), ("b_r1", C.c_int64), ("b_c1", C.c_int64),
        ("c_r0", C.c_int64), ("c_c0", C.c_int64), ("c_r1", C.c_int64), ("c_c1", C.c_int64),
        ("a_rows", C.c_int64), ("a_cols", C.c_int64), ("b_rows", C.c_int64), ("b_cols", C.c_int64),
        ("epilogue", C.c_int32),
        ("alpha", C.c_float),
        ("bias", _f32p), ("gamma", _f32p),
        ("add", C.c_void_p), ("ldadd", C.c_int64),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("tokens_per_frame", C.c_int32), ("patch_start", C.c_int32), ("grid_w", C.c_int32),
        ("apply_qknorm_rope", C.c_int32),
        ("qn_w", _f32p), ("qn_b", _f32p), ("kn_w", _f32p), ("kn_b", _f32p),
        ("qk_eps", C.c_float), ("q_scale", C.c_float),
        ("rope_cos", _f32p), ("rope_sin", _f32p),
        ("rope_positions", C.c_int32),
        ("n_patches", C.c_int32),
        ("pos_embed", _f32p),
        ("dropout", DropoutDesc),
        ("ln_stats", _f32p), ("ln_colsum", _f32p), ("ln_eps", C.c_float),
        ("stats_out", _f32p), ("stats_zero", _f32p), ("c_copy_bf16", C.c_void_p),
    ]


class BlockWeights(C.Structure):
    """Mirror of `vq_block_weights`."""

    _fields_ = [(n, C.c_void_p) for n in (
        "norm1_w", "norm1_b", "qkv_w", "qkv_b", "q_norm_w", "q_norm_b", "k_norm_w", "k_norm_b",
        "proj_w", "proj_b", "ls1", "norm2_w", "norm2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")] + [
        ("ln_folded", C.c_int32), ("qkv_colsum", C.c_void_p), ("fc1_colsum", C.c_void_p),
        ("qk_score_bound", C.c_float)]


class AggregatorWeights(C.Structure):
    """Mirror of `vq_aggregator_weights`."""

    _fields_ = [
        ("patch_w", C.c_void_p), ("patch_b", C.c_void_p),
        ("cls_token", C.c_void_p), ("pos_embed", C.c_void_p), ("register_tokens", C.c_void_p),
        ("stem_blocks", C.POINTER(BlockWeights)), ("n_stem", C.c_int32),
        ("stem_norm_w", C.c_void_p), ("stem_norm_b", C.c_void_p),
        ("camera_token", C.c_void_p), ("register_token", C.c_void_p),
        ("frame_blocks", C.POINTER(BlockWeights)), ("global_blocks", C.POINTER(BlockWeights)),
        ("depth", C.c_int32),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("rope_max_pos", C.c_int32),
    ]


class AggregatorCfg(C.Structure):
    """Mirror of `vq_aggregator_cfg`."""

    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32), ("patch", C.c_int32),
        ("D", C.c_int32), ("heads", C.c_int32), ("hidden", C.c_int32),
        ("n_special", C.c_int32), ("Kp", C.c_int32),
        ("stem_eps", C.c_float), ("block_eps", C.c_float), ("qk_eps", C.c_float),
        ("deferred_residual", C.c_int32), ("images_u8", C.c_int32),
    ]


class PerceiverCfg(C.Structure):
    """Mirror of `vq_perceiver_cfg`."""

    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("in_dim", C.c_int32), ("E", C.c_int32), ("heads", C.c_int32),
        ("Lq", C.c_int32), ("layers", C.c_int32), ("ffn", C.c_int32), ("out_dim", C.c_int32),
        ("ln_eps", C.c_float), ("out_dtype", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("dropout_offset", C.c_uint64),
    ]


_PLAYER_FIELDS = ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "norm1_w", "norm1_b",
                  "fc1_w", "fc1_b", "fc2_w", "fc2_b", "norm2_w", "norm2_b")


class PerceiverLayerWeights(C.Structure):
    """Mirror of `vq_perceiver_layer_weights`."""

    _fields_ = [(n, C.c_void_p) for n in _PLAYER_FIELDS]


class PerceiverWeights(C.Structure):
    """Mirror of `vq_perceiver_weights`."""

    _fields_ = [
        ("latents", C.c_void_p), ("in_proj_w", C.c_void_p), ("in_proj_b", C.c_void_p),
        ("layers", C.POINTER(PerceiverLayerWeights)),
        ("out_proj_w", C.c_void_p), ("out_proj_b", C.c_void_p),
    ]


class PerceiverLayerGrads(C.Structure):
    """Mirror of `vq_perceiver_layer_grads` (same field order as the weights)."""

    _fields_ = [(n, C.c_void_p) for n in _PLAYER_FIELDS]


class PerceiverGrads(C.Structure):
    """Mirror of `vq_perceiver_grads`."""

    _fields_ = [
        ("latents", C.c_void_p), ("in_proj_w", C.c_void_p), ("in_proj_b", C.c_void_p),
        ("layers", C.POINTER(PerceiverLayerGrads)),
        ("out_proj_w", C.c_void_p), ("out_proj_b", C.c_void_p),
        ("accumulate", C.c_int32), ("weight_grad_bf16", C.c_int32),
    ]


LayerDoneFn = C.CFUNCTYPE(None, C.c_int32, C.c_void_p)

_lib = None


def _declare(lib: C.CDLL) -> None:
    lib.vq_last_error.restype = C.c_char_p
    lib.vq_last_error.argtypes = []
    lib.vq_version.restype = C.c_int
    lib.vq_num_sms.restype = C.c_int
    lib.vq_gemm.restype = C.c_int
    lib.vq_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.vq_gemm_tune.restype = C.c_int
    lib.vq_gemm_tune.argtypes = [C.c_int32]
    lib.vq_attention_fwd.restype = C.c_int
    lib.vq_attention_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.vq_attention_tune.restype = C.c_int
    lib.vq_attention_tune.argtypes = [C.c_int32, C.c_int32]
    lib.vq_attention_fwd_bounded.restype = C.c_int
    lib.vq_attention_fwd_bounded.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    lib.vq_attention_tune_bounded.restype = C.c_int
    lib.vq_attention_tune_bounded.argtypes = [C.c_int32, C.c_int32]
    lib.vq_preprocess_workspace_bytes.restype = C.c_size_t
    lib.vq_preprocess_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.vq_aggregator_workspace_bytes.restype = C.c_size_t
    lib.vq_aggregator_workspace_bytes.argtypes = [C.POINTER(AggregatorCfg)]
    lib.vq_launch_count.restype = C.c_longlong
    lib.vq_launch_count.argtypes = []
    lib.vq_profile_enable.restype = C.c_int
    lib.vq_profile_enable.argtypes = [C.c_int]
    lib.vq_profile_dump.restype = C.c_longlong
    lib.vq_profile_dump.argtypes = [C.c_char_p, C.c_longlong]
    for name in ("vq_perceiver_saved_bytes", "vq_perceiver_workspace_bytes"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.POINTER(PerceiverCfg)]
    for name, argtypes in _EXTRA_DECLS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes


# name -> argtypes for the row/elementwise entry points (all return int status)
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
_EXTRA_DECLS: dict[str, list] = {
    "vq_im2col_patches": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "vq_im2col_patches_u8": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "vq_init_special_tokens": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "vq_layernorm": [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp],
    "vq_residual_layernorm": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp],
    "vq_row_stats_cast": [_vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "vq_stem_finalize": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp],
    "vq_gather_tokens": [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp],
    "vq_splice_overwrite": [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp],
    "vq_preprocess_geometry": [_i32, _i32, _i32, _vp],
    "vq_preprocess_taps": [_i32, _i32, _i32, _i32, _vp, _i64],
    "vq_preprocess_views": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, C.c_size_t, _vp],
    "vq_geom_head_forward": [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                             _i32, _i32, _i32, _vp],
    "vq_geom_head_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "vq_splice_features": [_vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp],
    "vq_splice_features_backward": [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp],
    "vq_find_first_token": [_vp, _i64, _i64, _vp, _vp],
    "vq_splice_insert": [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "vq_adamw_step": [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _vp],
    "vq_adamw_reduce_step": [_vp, _vp, _vp, C.POINTER(C.c_void_p), _i32, _i32, _vp, _i32, _i64, _i32, _f32, _f32, _f32,
                             _f32, _f32, _f32, _vp],
    "vq_softmax_fwd": [_vp, _vp, _i64, _i32, _i32, _i32, C.POINTER(DropoutDesc), _vp, _vp],
    "vq_softmax_bwd": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, C.POINTER(DropoutDesc), _vp],
    "vq_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, C.POINTER(DropoutDesc), _vp, _vp],
    "vq_perceiver_attention_supported": [_i32, _i32, _i32, _i32],
    "vq_perceiver_attention_tune": [_i32],
    "vq_perceiver_attention_fwd": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(DropoutDesc), _vp],
    "vq_perceiver_attention_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(DropoutDesc),
                                   _vp],
    "vq_dropout_mask": [_vp, _i64, _i32, _i32, C.POINTER(DropoutDesc), _vp],
    "vq_colsum": [_vp, _i32, _vp, _i64, _i32, _i64, _vp],
    "vq_cast_f32_bf16": [_vp, _vp, _i64, _vp],
    "vq_broadcast_rows": [_vp, _vp, _i64, _i32, _vp],
    "vq_reduce_batch": [_vp, _vp, _i64, _i32, _i32, _vp],
    "vq_perceiver_forward": [C.POINTER(PerceiverCfg), C.POINTER(PerceiverWeights), _vp, _vp, _vp, _vp, _vp],
    "vq_perceiver_backward": [C.POINTER(PerceiverCfg), C.POINTER(PerceiverWeights), _vp, _vp, _vp,
                              C.POINTER(PerceiverGrads), _vp, _vp, LayerDoneFn, _vp, _vp],
    "vq_aggregator_forward": [C.POINTER(AggregatorCfg), C.POINTER(AggregatorWeights), _vp, _vp, _i32,
                              C.POINTER(C.c_void_p), _vp, C.c_size_t, _vp],
}


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m vggt_qwen3_b200.csrc.build` "
                "(there is no CPU / PyTorch fallback for this path)"
            )
        handle = C.CDLL(os.fspath(LIB_PATH))
        _declare(handle)
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().vq_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")
