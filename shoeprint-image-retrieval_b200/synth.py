"""Seeded synthetic feature maps for tests and benchmarks (there is no network for datasets).

Gallery maps are print-like: low-pass noise plus fine texture, never exactly flat (a flat
window makes the reference divide round-off by ~0).  Probes are degraded crops of
their true gallery map so the true match separates from the rest.
"""

from __future__ import annotations

import numpy as np


def smooth_field(rng: np.random.Generator, c: int, h: int, w: int, passes: int = 2, gain: float = 6.0) -> np.ndarray:
    a = rng.standard_normal((c, h + 4, w + 4)).astype(np.float32)
    for _ in range(passes):
        a = (a + np.roll(a, 1, 1) + np.roll(a, 1, 2) + np.roll(a, -1, 1) + np.roll(a, -1, 2)) / 5
    a = a[:, 2:-2, 2:-2] * gain + 0.15 * rng.standard_normal((c, h, w)).astype(np.float32)
    return np.ascontiguousarray(a, dtype=np.float32)


def make_gallery(seed: int, g: int, c: int, h: int, w: int) -> list[np.ndarray]:
    rng = np.random.default_rng(seed)
    return [smooth_field(rng, c, h, w) for _ in range(g)]


def make_probes(seed: int, gallery: list[np.ndarray], q: int, min_frac: float = 1.0, noise: float = 0.3,
                pairs: list[int] | None = None) -> tuple[list[np.ndarray], list[int]]:
    """``q`` probes; probe i is a noisy crop of gallery[pairs[i]] whose height/width are drawn from
    ``[min_frac, 1] * gallery size`` (``min_frac=1`` -> all probes have the gallery's shape)."""
    rng = np.random.default_rng(seed)
    if pairs is None:
        pairs = [int(x) for x in rng.integers(0, len(gallery), size=q)]
    probes = []
    for i in range(q):
        gal = gallery[pairs[i]]
        _, h, w = gal.shape
        hq = int(rng.integers(max(6, int(h * min_frac)), h + 1))
        wq = int(rng.integers(max(6, int(w * min_frac)), w + 1))
        y0 = int(rng.integers(0, h - hq + 1))
        x0 = int(rng.integers(0, w - wq + 1))
        crop = gal[:, y0 : y0 + hq, x0 : x0 + wq]
        probes.append(np.ascontiguousarray(crop + noise * rng.standard_normal(crop.shape), dtype=np.float32))
    return probes, list(pairs)


def device_gallery(seed: int, g: int, c: int, h: int, w: int, device="cuda"):
    """Gallery maps generated on the device, ``[g, c, h, w]`` float32 (torch Philox normal, box
    smoothed).  Used where the set is too large to build on the host."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((g, c, h, w), dtype=torch.float32, device=device)
    step = max(1, min(g, (1 << 28) // (c * (h + 4) * (w + 4))))
    for s in range(0, g, step):
        n = min(step, g - s)
        a = torch.randn((n, c, h + 4, w + 4), generator=gen, device=device)
        a = torch.nn.functional.avg_pool2d(a, 3, stride=1, padding=1)
        a = torch.nn.functional.avg_pool2d(a, 3, stride=1, padding=1)
        out[s : s + n] = a[:, :, 2:-2, 2:-2] * 6.0 + 0.15 * torch.randn((n, c, h, w), generator=gen, device=device)
    return out


def device_probes(seed: int, gallery, q: int, noise: float = 0.3):
    """``q`` full-size noisy copies of randomly chosen device gallery maps + their indices."""
    import torch

    gen = torch.Generator(device=gallery.device)
    gen.manual_seed(seed)
    pairs = torch.randint(0, gallery.shape[0], (q,), generator=gen, device=gallery.device)
    probes = gallery[pairs] + noise * torch.randn((q, *gallery.shape[1:]), generator=gen, device=gallery.device)
    return probes.contiguous(), pairs.to(torch.int32)
