"""Oracle variant generation vs Pillow itself (the third-party code the reference calls)."""

import numpy as np
from PIL import Image

from oracle import variants


def _pil_rotate(m, angle):
    return np.array([np.array(Image.fromarray(ch).rotate(angle)) for ch in m])


def _pil_resize(m, s):
    out = []
    for ch in m:
        im = Image.fromarray(ch)
        out.append(np.array(im.resize((int(im.width * s), int(im.height * s)))))
    return np.array(out)


def test_rotate_bit_exact_random_cases():
    rng = np.random.default_rng(7)
    for _ in range(300):
        h, w = int(rng.integers(5, 70)), int(rng.integers(5, 40))
        angle = float(rng.choice([-30, -25, -20, -15, -10, -9, -5, -3, 3, 5, 9, 10, 15, 20, 25, 30, 45, 90, 180, 270, 0, 360, 7.5]))
        m = rng.standard_normal((2, h, w)).astype(np.float32)
        np.testing.assert_array_equal(variants.rotate_maps(m, angle), _pil_rotate(m, angle))


def test_resize_bit_exact_random_cases():
    rng = np.random.default_rng(8)
    for _ in range(150):
        h, w = int(rng.integers(6, 70)), int(rng.integers(6, 40))
        s = float(rng.choice([1.02, 1.04, 1.08, 0.8, 0.93, 1.25, 1.5, 2.0, 0.5]))
        m = rng.standard_normal((2, h, w)).astype(np.float32)
        want = _pil_resize(m, s)
        got = variants.resize_maps(m, s)
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)
