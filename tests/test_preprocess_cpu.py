"""The preprocessing oracle (oracle/preprocess_ref.py) against golden vectors produced by the
reference's own transform (tests/golden/make_golden_preprocess.py) — bit-exact — and, when torchvision is importable,
live against torchvision-on-PIL at the sizes the path uses (518 and the 448 of configs/stage1_3d.yaml)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import preprocess_ref as R

GOLD = Path(__file__).resolve().parent / "golden" / "preprocess_ref.npz"


def test_preprocess_oracle_matches_reference_golden_bit_exactly():
    z = np.load(GOLD)
    for i, (h, w, size) in enumerate(z["cases"]):
        img = z[f"in{i}"]
        assert img.shape == (h, w, 3)
        ours = R.preprocess_view(img, int(size))
        ref = z[f"out{i}"].astype(np.float32) / np.float32(255.0)
        assert ours.dtype == np.float32 and ours.shape == (3, size, size)
        assert np.array_equal(ours, ref), f"case {i}: {np.abs(ours - ref).max() * 255:.1f} LSB"


def test_geometry_helpers():
    assert R.resized_output_size(600, 800, 518) == (518, 690)      # shorter edge -> 518, int(518 * 800 / 600)
    assert R.resized_output_size(800, 600, 518) == (690, 518)
    assert R.resized_output_size(518, 518, 518) == (518, 518)
    assert R.center_crop_box(518, 690, 518) == (0, 86)
    assert R.center_crop_box(519, 518, 518) == (0, 0)               # python round(): half to even
    b, kk = R.precompute_coeffs(800, 690)
    assert (kk.sum(1) - (1 << R.PRECISION_BITS)).__abs__().max() <= kk.shape[1]   # taps sum to 1 in fixed point
    assert b[0, 0] == 0 and b[-1, 0] + b[-1, 1] == 800


def test_preprocess_oracle_matches_live_torchvision():
    tv = pytest.importorskip("torchvision")
    from PIL import Image
    from torchvision import transforms

    rng = np.random.default_rng(1)
    for h, w, size in [(480, 640, 448), (600, 800, 518), (700, 519, 518)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        t = transforms.Compose([transforms.Resize(size, interpolation=transforms.InterpolationMode.BICUBIC),
                                transforms.CenterCrop(size), transforms.ToTensor()])
        assert np.array_equal(R.preprocess_view(img, size), t(Image.fromarray(img)).numpy()), (h, w, size, tv.__version__)


def test_library_taps_and_geometry_match_the_oracle_bit_exactly():
    """The C++ side of the GPU preprocessing (csrc/preprocess.cu) computes its geometry and fixed-point taps on the
    host: both are checked here, without a GPU, against the oracle (which is pinned on the reference's transform)."""
    import ctypes

    from vggt_qwen3_b200 import _lib

    h = _lib.lib()
    for H, W, size in [(600, 800, 518), (480, 640, 448), (518, 518, 518), (1080, 1920, 518), (300, 200, 448),
                       (37, 53, 28), (700, 519, 518), (40, 60, 64), (519, 518, 518), (521, 518, 518)]:
        out6 = (ctypes.c_int32 * 6)()
        assert h.vq_preprocess_geometry(H, W, size, out6) == 0
        nh, nw = R.resized_output_size(H, W, size)
        top, left = R.center_crop_box(nh, nw, size)
        assert list(out6)[:4] == [nh, nw, top, left], (H, W, size, list(out6))
        for in_size, out_size, first, k_lib in ((W, nw, left, out6[4]), (H, nh, top, out6[5])):
            bounds, kk = R.precompute_coeffs(in_size, out_size)
            ksize = kk.shape[1]
            assert k_lib == ksize
            table = (ctypes.c_int32 * (size * (2 + ksize)))()
            assert h.vq_preprocess_taps(in_size, out_size, first, size, table, len(table)) == ksize
            t = np.frombuffer(table, dtype=np.int32).reshape(size, 2 + ksize)
            assert np.array_equal(t[:, 0], bounds[first:first + size, 0])
            assert np.array_equal(t[:, 1], bounds[first:first + size, 1])
            assert np.array_equal(t[:, 2:], kk[first:first + size])
    assert h.vq_preprocess_workspace_bytes(0, 10, 10, 4) == 0 and b"geometry" in h.vq_last_error()
    assert h.vq_preprocess_workspace_bytes(8, 600, 800, 518) >= 8 * 600 * 518 * 3
