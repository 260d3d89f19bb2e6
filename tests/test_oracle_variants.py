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


def test_loader_lanczos_restatement_is_bit_exact_vs_pillow():
    """oracle.loader.resize_lanczos_u8 == Image.resize(size, LANCZOS) on 8-bit images (dataloader.py:231-237): gray and RGB,
    enlargement and reduction, one or both axes."""
    import numpy as np
    from PIL import Image

    from oracle.loader import resize_lanczos_u8

    rng = np.random.default_rng(0)
    for t in range(40):
        h, w = int(rng.integers(5, 90)), int(rng.integers(5, 90))
        s = float(rng.uniform(0.3, 1.6))
        h2 = max(1, int(h * s)) if t % 5 else h
        w2 = max(1, int(w * (s if t % 3 else rng.uniform(0.3, 1.6))))
        img = rng.integers(0, 256, size=(h, w) if t % 4 else (h, w, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((w2, h2), Image.Resampling.LANCZOS))
        np.testing.assert_array_equal(resize_lanczos_u8(img, h2, w2), want)
