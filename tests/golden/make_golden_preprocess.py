"""Golden vectors for the image preprocessing: outputs of the REFERENCE's own transform
(src/dataio/collate_multiview.py:12-19, imported from /root/reference) on seeded uint8 images, PIL 12.2 /
torchvision 0.26.  Run here (CPU):  python tests/golden/make_golden_preprocess.py
The outputs are stored as uint8 (ToTensor output is exactly u8 / 255)."""
import sys
from pathlib import Path

import numpy as np
from PIL import Image

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from src.dataio.collate_multiview import build_default_transform  # noqa: E402

CASES = [(37, 53, 28), (120, 90, 64), (100, 100, 64), (40, 60, 64), (151, 97, 56), (64, 200, 42)]  # (h, w, image_size)

if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for i, (h, w, size) in enumerate(CASES):
        # half noise, half smooth gradient: exercises the clamp at 0 / 255 and the negative bicubic lobes
        yy, xx = np.mgrid[0:h, 0:w]
        smooth = np.stack([(yy * 255 // max(h - 1, 1)), (xx * 255 // max(w - 1, 1)), ((yy + xx) % 256)], -1)
        noise = rng.integers(0, 256, (h, w, 3))
        img = np.where((yy // 8 + xx // 8)[..., None] % 2 == 0, smooth, noise).astype(np.uint8)
        ref = build_default_transform(size)(Image.fromarray(img)).numpy()
        u8 = np.rint(ref * 255.0).astype(np.uint8)
        assert np.array_equal(u8.astype(np.float32) / np.float32(255.0), ref)
        out[f"in{i}"], out[f"out{i}"] = img, u8
    path = Path(__file__).resolve().parent / "preprocess_ref.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes")
