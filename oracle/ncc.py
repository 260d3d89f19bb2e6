"""Oracle: normalised cross-correlation of one probe feature map against one gallery map.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``src/shoeprint_image_retrieval/similarity.py:26-108`` of the reference
(``normxcorr`` lines 26-72, ``get_similarity`` lines 75-108).  The reference calls
``scipy.signal.convolve(..., mode="same")`` (third-party, pinned scipy 1.14.0 in the
reference's ``uv.lock``) three times per channel; the published algorithm of that call is a
zero-padded linear convolution whose output keeps the size of the FIRST argument and is
centred on the full result, i.e. ``same[y] = full[y + (Hm - 1) // 2]``.  With the flipped
template the reference passes, this is the correlation

    num[y, x] = sum_{u < Hm, v < Wm} t[u, v] * g[y + u - Hm // 2, x + v - Wm // 2]

with zeros outside ``g``.  Two implementations are kept:

``ncc_surface_direct``  float64, evaluates the windowed definition literally (ground truth).
``ncc_surface_fft``     follows the reference's arithmetic: float32 numerator through an FFT
                        convolution, float64 box sums through FFT convolutions with a float64
                        ones kernel.  This is the one timed as the CPU baseline ("port").
"""

from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from scipy import fft as _fft

__all__ = [
    "crop_edges",
    "ncc_surface_direct",
    "ncc_surface_fft",
    "get_similarity",
    "window_denominator",
]

_EDGE = 2  # similarity.py:92-93 crops two cells per edge from both maps


def crop_edges(maps: np.ndarray) -> np.ndarray:
    """``maps[:, 2:-2, 2:-2]`` (similarity.py:92-93)."""
    return maps[:, _EDGE:-_EDGE, _EDGE:-_EDGE]


def _pad_for_same(g: np.ndarray, hm: int, wm: int) -> np.ndarray:
    """Zero-pad ``g`` so a *valid* window scan reproduces scipy's ``mode="same"``.

    Anchor (Hm // 2, Wm // 2): top/left pad = size // 2, bottom/right = size - 1 - size // 2.
    """
    top, left = hm // 2, wm // 2
    return np.pad(g, ((top, hm - 1 - top), (left, wm - 1 - left)))


def window_denominator(g0: np.ndarray, hm: int, wm: int) -> np.ndarray:
    """Local energy term D of similarity.py:57-65 for an already zero-meaned image, in float64.

    D[y,x] = S2 - S1^2 / (Hm*Wm) over the (zero padded) template-sized window, clamped at 0.
    The divisor is the full template size everywhere, borders included.
    """
    gp = _pad_for_same(g0.astype(np.float64), hm, wm)
    # summed-area tables (exact enough in float64: |values| * count << 2^53)
    def box(a: np.ndarray) -> np.ndarray:
        sat = np.zeros((a.shape[0] + 1, a.shape[1] + 1))
        sat[1:, 1:] = a.cumsum(0).cumsum(1)
        return sat[hm:, wm:] - sat[:-hm, wm:] - sat[hm:, :-wm] + sat[:-hm, :-wm]

    s1 = box(gp)
    s2 = box(gp * gp)
    d = s2 - s1 * s1 / float(hm * wm)
    d[d < 0] = 0.0
    return d


def ncc_surface_direct(template: np.ndarray, image: np.ndarray) -> np.ndarray:
    """NCC surface, gallery sized, float64, by the windowed definition (ground truth)."""
    t = template.astype(np.float32)
    g = image.astype(np.float32)
    t0 = (t - np.mean(t)).astype(np.float64)  # similarity.py:48 (float32 mean, as numpy does)
    g0 = g - np.mean(g)  # similarity.py:49
    hm, wm = t0.shape
    win = sliding_window_view(_pad_for_same(g0.astype(np.float64), hm, wm), (hm, wm))
    num = np.einsum("yxuv,uv->yx", win, t0, optimize=True)
    d = window_denominator(g0, hm, wm)
    e = float(np.sum(t0 * t0))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / np.sqrt(d * e)
    out[~np.isfinite(out)] = 0.0  # similarity.py:70
    return out


def _same_conv_fft(a: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Linear convolution of ``a`` with ``k`` cropped like scipy ``mode="same"`` (size of ``a``)."""
    dt = np.result_type(a.dtype, k.dtype)
    fh = _fft.next_fast_len(a.shape[0] + k.shape[0] - 1, real=True)
    fw = _fft.next_fast_len(a.shape[1] + k.shape[1] - 1, real=True)
    spec = _fft.rfft2(a.astype(dt), (fh, fw)) * _fft.rfft2(k.astype(dt), (fh, fw))
    full = _fft.irfft2(spec, (fh, fw))
    y0, x0 = (k.shape[0] - 1) // 2, (k.shape[1] - 1) // 2
    return full[y0 : y0 + a.shape[0], x0 : x0 + a.shape[1]].astype(dt, copy=False)


def ncc_surface_fft(template: np.ndarray, image: np.ndarray) -> np.ndarray:
    """NCC surface with the reference's arithmetic (similarity.py:48-70): three "same"
    convolutions per call, float32 numerator, float64 box sums."""
    t0 = template - np.mean(template)
    g0 = image - np.mean(image)
    ones = np.ones(t0.shape)  # float64 on purpose: similarity.py:50
    num = _same_conv_fft(g0, t0[::-1, ::-1])
    s2 = _same_conv_fft(np.square(g0), ones)
    s1 = _same_conv_fft(g0, ones)
    d = s2 - np.square(s1) / t0.size
    d[d < 0] = 0
    e = np.sum(np.square(t0))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / np.sqrt(d * e)
    out[~np.isfinite(out)] = 0
    return out


def get_similarity(shoemark: np.ndarray, shoeprint: np.ndarray, method: str = "direct") -> float:
    """Score of one probe(-variant) map against one gallery map (similarity.py:75-108):
    crop 2 cells/edge, per-channel NCC, sum over channels, max over positions, / C."""
    mark = crop_edges(shoemark)
    prnt = crop_edges(shoeprint)
    surface = ncc_surface_direct if method == "direct" else ncc_surface_fft
    total = np.zeros(prnt.shape[1:], dtype=np.float64)
    for c in range(mark.shape[0]):
        total += surface(mark[c], prnt[c])
    return float(total.max() / mark.shape[0])
