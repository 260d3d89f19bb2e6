"""CPU restatement of the loader's resize (test infrastructure only): ``Image.resize(size, LANCZOS)`` on 8-bit images as
the reference's loader calls it (``dataloader.py:231-237``).  The arithmetic lives in Pillow's ``Resample.c``
(``precompute_coeffs``, ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``; pinned 10.2.0,
installed 12.2.0), which is not part of the reference tree; ``tests/test_oracle_variants.py`` pins this restatement bit for
bit against the installed Pillow."""

from __future__ import annotations

import math

import numpy as np

_BITS = 22  # PRECISION_BITS = 32 - 8 - 2


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        def sinc(v: float) -> float:
            if v == 0.0:
                return 1.0
            v *= math.pi
            return math.sin(v) / v
        return sinc(x) * sinc(x / 3)
    return 0.0


def _coeffs(n_in: int, n_out: int) -> tuple[list[int], list[list[int]]]:
    scale = n_in / n_out
    fscale = max(scale, 1.0)
    support = 3.0 * fscale
    ss = 1.0 / fscale
    xmins, taps = [], []
    for xx in range(n_out):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), n_in)
        k = [_lanczos((i + lo - center + 0.5) * ss) for i in range(hi - lo)]
        ww = 0.0
        for v in k:
            ww += v
        if ww != 0.0:
            k = [v / ww for v in k]
        xmins.append(lo)
        taps.append([int(-0.5 + v * (1 << _BITS)) if v < 0 else int(0.5 + v * (1 << _BITS)) for v in k])
    return xmins, taps


def _pass(img: np.ndarray, n_out: int, axis: int) -> np.ndarray:
    """One 8-bit pass along ``axis`` (0 = rows / vertical, 1 = columns / horizontal) of ``[h, w, ch]``."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    xmins, taps = _coeffs(src.shape[0], n_out)
    out = np.empty((n_out, *src.shape[1:]), dtype=np.uint8)
    for o in range(n_out):
        acc = np.full(src.shape[1:], 1 << (_BITS - 1), dtype=np.int64)
        for j, k in enumerate(taps[o]):
            acc += src[xmins[o] + j] * k
        out[o] = np.clip(acc >> _BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_lanczos_u8(img: np.ndarray, h2: int, w2: int) -> np.ndarray:
    """``np.array(Image.fromarray(img).resize((w2, h2), Image.Resampling.LANCZOS))`` for uint8 ``[h, w]`` or ``[h, w, 3]``."""
    a = img if img.ndim == 3 else img[:, :, None]
    h, w = a.shape[:2]
    if w2 != w:
        a = _pass(a, w2, 1)  # horizontal pass first (ImagingResample)
    if h2 != h:
        a = _pass(a, h2, 0)
    a = np.ascontiguousarray(a)
    return a if img.ndim == 3 else a[:, :, 0]
