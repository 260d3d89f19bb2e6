"""Oracle CLAHE restatement vs OpenCV itself (the third-party code the reference calls)."""

import cv2
import numpy as np

from oracle import clahe as oclahe


def test_clahe_bit_exact_vs_opencv():
    rng = np.random.default_rng(5)
    cases = [(64, 64), (470, 162), (800, 300), (586, 270), (123, 77), (40, 200), (17, 33)]
    for h, w in cases:
        base = rng.integers(0, 256, size=(h // 7 + 1, w // 7 + 1)).astype(np.float32)
        img = np.kron(base, np.ones((7, 7), np.float32))[:h, :w]
        img = np.clip(img * 0.7 + rng.normal(30, 20, img.shape), 0, 255).astype(np.uint8)
        for clip, tiles in [(2.0, (8, 8)), (4.0, (4, 6)), (0.0, (8, 8)), (40.0, (2, 2))]:
            want = cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles).apply(img)
            got = oclahe.clahe(img, clip, tiles)
            np.testing.assert_array_equal(got, want, err_msg=f"{h}x{w} clip {clip} tiles {tiles}")
