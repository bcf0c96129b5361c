"""GPU parity of the uint8 image preprocessing (csrc/preprocess.cu) through the C ABI against the
oracle restatement of the reference transform (src/dataio/collate_multiview.py:12-19): integer / byte work, so the
bar is BIT-EXACT; plus the committed goldens made with the reference's own transform."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import preprocess_ref as R

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "preprocess_ref.npz"


@pytest.mark.parametrize("frames,H,W,size", [
    (2, 480, 640, 448), (1, 600, 800, 518), (3, 37, 53, 28), (1, 518, 518, 518), (2, 40, 60, 64), (1, 700, 519, 518),
    (1, 300, 200, 448),
])
def test_preprocess_views_bit_exact_against_oracle(frames, H, W, size):
    from vggt_qwen3_b200 import ops

    rng = np.random.default_rng(H * 1000 + W)
    views = rng.integers(0, 256, (frames, H, W, 3), dtype=np.uint8)
    views[0, : H // 2] = (np.arange(W)[None, :, None] * 255 // max(W - 1, 1)).astype(np.uint8)   # smooth half
    out = ops.preprocess_views(torch.from_numpy(views).cuda(), size)
    torch.cuda.synchronize()
    assert out.shape == (frames, 3, size, size) and out.dtype == torch.float32
    ref = np.stack([R.preprocess_view(views[f], size) for f in range(frames)])
    assert np.array_equal(out.cpu().numpy(), ref), f"max diff {np.abs(out.cpu().numpy() - ref).max() * 255:.2f} LSB"


def test_preprocess_views_reference_goldens_and_batch_shape():
    from vggt_qwen3_b200 import ops

    z = np.load(GOLD)
    for i, (h, w, size) in enumerate(z["cases"]):
        out = ops.preprocess_views(torch.from_numpy(z[f"in{i}"]).cuda(), int(size))
        ref = z[f"out{i}"].astype(np.float32) / np.float32(255.0)
        assert np.array_equal(out.cpu().numpy(), ref), i
    # [B, V, H, W, 3] -> [B, V, 3, S, S]: the layout encode_images takes
    views = torch.randint(0, 256, (2, 3, 90, 120, 3), dtype=torch.uint8)
    out = ops.preprocess_views(views.cuda(), 56)
    assert out.shape == (2, 3, 3, 56, 56)
    one = ops.preprocess_views(views[1, 2].cuda(), 56)
    assert torch.equal(out[1, 2], one)
