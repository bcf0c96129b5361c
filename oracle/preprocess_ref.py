"""ORACLE (test infrastructure only — never imported by the product path): CPU restatement of the image
preprocessing that feeds the visual-encode path.

Reference: src/dataio/collate_multiview.py:12-19 (`build_default_transform`) applied per view at :38-40:
    transforms.Resize(image_size, interpolation=BICUBIC)   shorter edge -> image_size, aspect kept (PIL image input,
                                                           so torchvision calls PIL.Image.resize(..., BICUBIC))
    transforms.CenterCrop(image_size)
    transforms.ToTensor()                                  uint8 HWC -> float32 CHW in [0, 1]
The arithmetic lives in two third-party dependencies that are not vendored in the reference and are restated here from
their published algorithms: torchvision 0.26 (`_compute_resized_output_size`, `center_crop`: integer geometry) and
Pillow 12.2 `src/libImaging/Resample.c` (separable antialiased convolution, bicubic a = -0.5, 22-bit fixed-point
coefficients, uint8 rounding after EACH pass, horizontal pass first).  Pinned bit-exactly against
torchvision-on-PIL outputs: tests/golden/preprocess_ref.npz (made by tests/golden/make_golden_preprocess.py) and live
in tests/test_preprocess_cpu.py when torchvision is importable.  Integer / byte work: the parity bar is bit-exact.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2   # Resample.c: coefficients are scaled by 2^22


def resized_output_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision.transforms.functional._compute_resized_output_size for an int `size` (no max_size): the shorter
    edge becomes `size`, the longer one int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)   # (new_h, new_w)


def center_crop_box(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision center_crop offsets (top, left) for an image at least `size` in both dimensions."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box (in0 = 0, in1 = in_size):
    per output index the first input index, the tap count and the int32 fixed-point taps."""
    support0 = 2.0
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        k = np.array([_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)], dtype=np.float64)
        ww = k.sum()
        if ww != 0.0:
            k = k / ww
        fixed = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS)))
        kk[xx, :xmax] = fixed.astype(np.int64)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One pass of ImagingResample{Horizontal,Vertical}_8bpc over `axis` of a uint8 HWC array."""
    in_size = img.shape[axis]
    bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    half = 1 << (PRECISION_BITS - 1)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = half + np.tensordot(kk[xx, :n], src[x0:x0 + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """PIL.Image.resize((new_w, new_h), BICUBIC) on a uint8 HWC array: horizontal pass, then vertical pass, each
    skipped when that dimension does not change (Resample.c: need_horizontal / need_vertical)."""
    out = img
    if new_w != img.shape[1]:
        out = _resample_axis(out, new_w, axis=1)
    if new_h != img.shape[0]:
        out = _resample_axis(out, new_h, axis=0)
    return out


def preprocess_view(img_u8_hwc: np.ndarray, image_size: int) -> np.ndarray:
    """build_default_transform(image_size) on one RGB view: float32 [3, image_size, image_size] in [0, 1]."""
    h, w = img_u8_hwc.shape[:2]
    nh, nw = resized_output_size(h, w, image_size)
    r = pil_resize_bicubic(img_u8_hwc, nh, nw)
    top, left = center_crop_box(nh, nw, image_size)
    c = r[top:top + image_size, left:left + image_size]
    return np.ascontiguousarray(c.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
