"""Oracle: rotation / scale variants of probe feature maps.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``_apply_transformations`` (``src/shoeprint_image_retrieval/similarity.py:230-284``)
and the variant-set construction of ``_comparison_worker`` (``similarity.py:321-353``).
The per-channel arithmetic is Pillow's (third-party, pinned 10.2.0 in ``uv.lock``; 12.2.0
installed here), reached through ``Image.fromarray(float32)`` -> mode "F":

* ``Image.rotate(angle)``: nearest neighbour, same size, zero fill, centre (w/2, h/2).  The
  affine coefficients are rounded to 15 decimals, converted to 16.16 fixed point and walked
  incrementally; the source index is ``coord >> 16``.  0 deg is a copy, 180 deg an exact flip,
  90/270 a transpose only for square maps.
* ``Image.resize((int(w*s), int(h*s)))``: default filter for mode "F" is bicubic (a = -0.5,
  support 2), separable, horizontal pass first into a float32 intermediate; each pass
  accumulates ``float32 pixel * double weight`` in double and casts to float32.

Both restatements are checked bit-for-bit against Pillow in ``tests/test_oracle_variants.py``
and against reference-generated goldens.
"""

from __future__ import annotations

import math

import numpy as np

__all__ = [
    "rotate_index_map",
    "rotate_maps",
    "bicubic_coeffs",
    "resize_maps",
    "scaled_size",
    "build_variant_lists",
    "variant_plan",
]


# ---------------------------------------------------------------------------- rotation

def _fix16(v: float) -> int:
    return int(math.floor(v * 65536.0 + 0.5))


def rotate_index_map(h: int, w: int, angle: float) -> np.ndarray:
    """Source flat index (y*w+x) for every output cell of an ``h x w`` map rotated by ``angle``
    degrees counter-clockwise, or -1 where the output is zero fill."""
    angle = angle % 360.0
    yy, xx = np.mgrid[0:h, 0:w]
    if angle == 0:
        return (yy * w + xx).astype(np.int64)
    if angle == 180:
        return ((h - 1 - yy) * w + (w - 1 - xx)).astype(np.int64)
    if angle in (90, 270) and w == h:
        if angle == 90:  # Image.Transpose.ROTATE_90: out[y][x] = in[x][w-1-y]
            return (xx * w + (w - 1 - yy)).astype(np.int64)
        return ((h - 1 - xx) * w + yy).astype(np.int64)
    r = -math.radians(angle)
    m0 = round(math.cos(r), 15)
    m1 = round(math.sin(r), 15)
    m3 = round(-math.sin(r), 15)
    m4 = round(math.cos(r), 15)
    cx, cy = w / 2.0, h / 2.0
    m2 = m0 * (-cx) + m1 * (-cy) + 0.0 + cx
    m5 = m3 * (-cx) + m4 * (-cy) + 0.0 + cy
    a0, a1, a3, a4 = _fix16(m0), _fix16(m1), _fix16(m3), _fix16(m4)
    a2 = _fix16(m2 + m0 * 0.5 + m1 * 0.5)
    a5 = _fix16(m5 + m3 * 0.5 + m4 * 0.5)
    xs = (a2 + a1 * yy.astype(np.int64) + a0 * xx.astype(np.int64)) >> 16
    ys = (a5 + a4 * yy.astype(np.int64) + a3 * xx.astype(np.int64)) >> 16
    ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
    return np.where(ok, ys * w + xs, -1).astype(np.int64)


def rotate_maps(maps: np.ndarray, angle: float) -> np.ndarray:
    """Rotate every channel of ``maps [C,h,w]`` (float32) like ``Image.rotate(angle)``."""
    c, h, w = maps.shape
    idx = rotate_index_map(h, w, angle).reshape(-1)
    flat = maps.reshape(c, h * w)
    out = np.where(idx[None, :] >= 0, flat[:, np.clip(idx, 0, None)], np.float32(0))
    return out.reshape(c, h, w).astype(np.float32)


# ---------------------------------------------------------------------------- bicubic resize

def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_coeffs(n_in: int, n_out: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per output index: first source index, tap count, normalised double weights [n_out, ksize]."""
    scale = n_in / n_out
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(n_out, dtype=np.int64)
    cnt = np.zeros(n_out, dtype=np.int64)
    kk = np.zeros((n_out, ksize), dtype=np.float64)
    inv = 1.0 / fscale
    for xx in range(n_out):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, n_in)
        n = hi - lo
        ww = 0.0
        for i in range(n):
            wgt = _bicubic((i + lo - center + 0.5) * inv)
            kk[xx, i] = wgt
            ww += wgt
        if ww != 0.0:
            for i in range(n):
                kk[xx, i] /= ww
        xmin[xx] = lo
        cnt[xx] = n
    return xmin, cnt, kk


def _resample_axis(img: np.ndarray, n_out: int, axis: int) -> np.ndarray:
    """One separable pass along ``axis`` of ``img [C,h,w]`` float32 -> float32."""
    n_in = img.shape[axis]
    xmin, cnt, kk = bicubic_coeffs(n_in, n_out)
    src = np.moveaxis(img, axis, -1).astype(np.float64)
    out = np.zeros(src.shape[:-1] + (n_out,), dtype=np.float64)
    for xx in range(n_out):
        acc = np.zeros(src.shape[:-1], dtype=np.float64)
        for i in range(int(cnt[xx])):  # sequential double accumulation, Pillow's order
            acc = acc + src[..., xmin[xx] + i] * kk[xx, i]
        out[..., xx] = acc
    return np.moveaxis(out.astype(np.float32), -1, axis)


def scaled_size(h: int, w: int, s: float) -> tuple[int, int]:
    """``(int(h*s), int(w*s))`` exactly as similarity.py:269-274 computes it."""
    return int(h * s), int(w * s)


def resize_maps(maps: np.ndarray, s: float) -> np.ndarray:
    """Resize every channel of ``maps [C,h,w]`` like ``Image.resize((int(w*s), int(h*s)))``."""
    _, h, w = maps.shape
    h2, w2 = scaled_size(h, w, s)
    out = maps.astype(np.float32)
    if (h2, w2) == (h, w):
        return out.copy()
    if w2 != w:  # horizontal pass first, skipped when the width is unchanged
        out = _resample_axis(out, w2, 2)
    if h2 != h:
        out = _resample_axis(out, h2, 1)
    return out


# ---------------------------------------------------------------------------- variant set

def variant_plan(rotations, scales) -> list[tuple[float | None, float | None]]:
    """The (rotation, scale) recipe of every variant, in the order the reference scores them.

    similarity.py:321-353 + 282: none -> [id]; rotations only -> [id, r1..rR]; scales only ->
    [id, s1..sS]; both -> [id] + [scale_s(v) for v in (id, r1..rR) for s in scales]
    (the rotated-only lists are dropped: SURVEY.md Appendix D1), i.e. 1 + (R+1)*S variants.
    """
    if rotations is None and scales is None:
        return [(None, None)]
    if scales is None:
        return [(None, None)] + [(r, None) for r in rotations]
    if rotations is None:
        return [(None, None)] + [(None, s) for s in scales]
    plan: list[tuple[float | None, float | None]] = [(None, None)]
    for r in [None, *rotations]:
        for s in scales:
            plan.append((r, s))
    return plan


def apply_variant(maps: np.ndarray, rot, scale) -> np.ndarray:
    out = maps
    if rot is not None:
        out = rotate_maps(out, rot)
    if scale is not None:
        out = resize_maps(out, scale)
    return out


def build_variant_lists(probe_maps: list[np.ndarray], rotations, scales) -> list[list[np.ndarray]]:
    """List (per variant) of lists (per probe) of ``[C,h',w']`` float32 maps."""
    return [[apply_variant(m, r, s) for m in probe_maps] for (r, s) in variant_plan(rotations, scales)]
