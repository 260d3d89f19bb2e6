"""Oracle: all-pairs scoring, ranking and S-scores.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``_comparison_worker`` (``similarity.py:287-375``: zero-initialised float32 score
matrix, running max over variants, lines 355-367), ``_get_rank`` (``similarity.py:378-386``)
and ``cmp`` / ``cmp_all`` (``parse_results.py:4-35``).
"""

from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np
from scipy import fft as _fft

from .ncc import crop_edges, get_similarity, window_denominator
from .variants import build_variant_lists

__all__ = [
    "score_matrix",
    "score_matrix_fast",
    "ranks_from_scores",
    "rank_interval",
    "s_score",
    "s_scores",
    "compare_maps_oracle",
    "pair_variants_per_second",
]


def score_matrix(probe_maps, gallery_maps, rotations=None, scales=None, method="direct") -> np.ndarray:
    """float32 [Q,G]: max over the reference's variant set of ``get_similarity``, floored at 0."""
    variant_lists = build_variant_lists(list(probe_maps), rotations, scales)
    best = np.zeros((len(probe_maps), len(gallery_maps)), dtype=np.float32)  # similarity.py:355
    for probes_v in variant_lists:
        for qi, mark in enumerate(probes_v):
            for gi, prnt in enumerate(gallery_maps):
                sim = get_similarity(mark, prnt, method=method)
                if sim > best[qi, gi]:  # similarity.py:365-367
                    best[qi, gi] = sim
    return best


class _GalleryCache:
    """Per-gallery float64 spectra and window denominators, shared by all probe variants."""

    def __init__(self, gallery_maps):
        self.g0 = []
        for g in gallery_maps:
            gc = crop_edges(np.asarray(g, dtype=np.float32))
            gc = gc - gc.mean(axis=(1, 2), keepdims=True, dtype=np.float32)
            self.g0.append(gc.astype(np.float64))
        self._den: dict = {}
        self._spec: dict = {}

    def inv_sqrt_den(self, gi: int, hm: int, wm: int) -> np.ndarray:
        key = (gi, hm, wm)
        if key not in self._den:
            g0 = self.g0[gi]
            d = np.stack([window_denominator(g0[c], hm, wm) for c in range(g0.shape[0])])
            with np.errstate(divide="ignore"):
                r = 1.0 / np.sqrt(d)
            r[~np.isfinite(r)] = 0.0
            self._den[key] = r
        return self._den[key]

    def spectrum(self, gi: int, fh: int, fw: int) -> np.ndarray:
        key = (gi, fh, fw)
        if key not in self._spec:
            self._spec[key] = _fft.rfft2(self.g0[gi], (fh, fw), axes=(1, 2))
        return self._spec[key]


def score_matrix_fast(probe_maps, gallery_maps, rotations=None, scales=None) -> np.ndarray:
    """Same quantity as :func:`score_matrix` in float64 throughout, channels batched through
    one FFT per pair-variant and the gallery-only terms cached.  Used where the literal
    triple loop would take minutes; checked against it in ``tests/test_oracle_golden.py``."""
    variant_lists = build_variant_lists(list(probe_maps), rotations, scales)
    cache = _GalleryCache(gallery_maps)
    best = np.zeros((len(probe_maps), len(gallery_maps)), dtype=np.float32)
    for probes_v in variant_lists:
        for qi, mark in enumerate(probes_v):
            t = crop_edges(np.asarray(mark, dtype=np.float32))
            t = (t - t.mean(axis=(1, 2), keepdims=True, dtype=np.float32)).astype(np.float64)
            c, hm, wm = t.shape
            e = (t * t).sum(axis=(1, 2))
            with np.errstate(divide="ignore"):
                inv_e = np.where(e > 0, 1.0 / np.sqrt(e), 0.0)
            tn = t * inv_e[:, None, None]
            for gi in range(len(gallery_maps)):
                g0 = cache.g0[gi]
                hp, wp = g0.shape[1:]
                fh = _fft.next_fast_len(hp + hm - 1, real=True)
                fw = _fft.next_fast_len(wp + wm - 1, real=True)
                spec = cache.spectrum(gi, fh, fw) * _fft.rfft2(tn[:, ::-1, ::-1], (fh, fw), axes=(1, 2))
                full = _fft.irfft2(spec, (fh, fw), axes=(1, 2))
                y0, x0 = (hm - 1) // 2, (wm - 1) // 2
                num = full[:, y0 : y0 + hp, x0 : x0 + wp]
                surf = (num * cache.inv_sqrt_den(gi, hm, wm)).sum(axis=0)
                sim = surf.max() / c
                if sim > best[qi, gi]:
                    best[qi, gi] = sim
    return best


def ranks_from_scores(scores: np.ndarray, matching_pairs) -> np.ndarray:
    """1-based rank of the true match in the descending argsort of each row
    (similarity.py:381-386); ties resolved the way numpy's argsort + flip resolves them."""
    out = np.zeros(scores.shape[0], dtype=np.int32)
    for qi in range(scores.shape[0]):
        order = np.flip(np.argsort(scores[qi]))
        out[qi] = int(np.where(order == matching_pairs[qi])[0][0]) + 1
    return out


def rank_interval(row: np.ndarray, true_idx: int, rel_tol: float = 1e-4) -> tuple[int, int]:
    """Tolerance-aware rank bounds: with tau = rel_tol * s_true any
    rank in [1 + #{s > s_true + tau}, 1 + #{s >= s_true - tau, g != true}] is acceptable."""
    s_true = float(row[true_idx])
    tau = rel_tol * abs(s_true) + 1e-7
    others = np.delete(np.asarray(row, dtype=np.float64), true_idx)
    return 1 + int((others > s_true + tau).sum()), 1 + int((others >= s_true - tau).sum())


def s_score(rankings, p: int, total_shoeprints: int, total_shoemarks: int) -> float:
    """parse_results.py:4-24: share of probes whose rank is within p percent of the gallery."""
    limit = (p * total_shoeprints) / 100
    return sum(1 for r in rankings if r <= limit) / total_shoemarks


def s_scores(rankings, total_shoeprints: int, total_shoemarks: int) -> dict[str, float]:
    """parse_results.py:27-35 as values (percent) instead of a printed line."""
    return {f"S{p}": s_score(rankings, p, total_shoeprints, total_shoemarks) * 100 for p in (1, 5, 10, 15, 20)}


def compare_maps_oracle(probe_maps, gallery_maps, matching_pairs, rotations=None, scales=None, method="fast"):
    """(ranks int32 [Q], scores float32 [Q,G]) -- what the reference's workers write."""
    if method == "fast":
        scores = score_matrix_fast(probe_maps, gallery_maps, rotations, scales)
    else:
        scores = score_matrix(probe_maps, gallery_maps, rotations, scales, method=method)
    return ranks_from_scores(scores, matching_pairs), scores


# ---------------------------------------------------------------------------- CPU baseline

def _baseline_worker(args):
    probes, gallery, rotations, scales = args
    return score_matrix(probes, gallery, rotations, scales, method="fft")


def pair_variants_per_second(probe_maps, gallery_maps, rotations, scales, n_procs: int | None = None):
    """Time the reference-faithful scorer (``method="fft"``: three FFT convolutions per channel
    per pair, like similarity.py:55-59) over ``n_procs`` forked workers splitting the probes
    the way ``compare_maps`` does (similarity.py:146-157).  Returns (scores, seconds, n_procs)."""
    import time

    n_procs = n_procs or os.cpu_count() or 1
    n_procs = max(1, min(n_procs, len(probe_maps)))
    base, extra = divmod(len(probe_maps), n_procs)
    chunks, start = [], 0
    for i in range(n_procs):
        end = start + base + (1 if i < extra else 0)
        chunks.append((list(probe_maps[start:end]), list(gallery_maps), rotations, scales))
        start = end
    t0 = time.perf_counter()
    if n_procs == 1:
        parts = [_baseline_worker(chunks[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            parts = pool.map(_baseline_worker, chunks)
    dt = time.perf_counter() - t0
    return np.concatenate(parts, axis=0), dt, n_procs
