"""Ranks -> S-scores (cumulative match characteristic), reference ``parse_results.py:4-35``."""

from __future__ import annotations

S_PERCENTAGES = (1, 5, 10, 15, 20)


def cmp(rankings: list[int], p: int, total_shoeprints: int, total_shoemarks: int) -> float:
    """S_p: share of all shoemarks whose true match is ranked within the best ``p`` percent of the
    gallery (``rank <= p * total_shoeprints / 100``), divided by ``total_shoemarks``
    (``parse_results.py:14-24``; the caller passes whole-dataset totals, ``run.py:30-34``)."""
    cutoff = (p * total_shoeprints) / 100
    hits = sum(1 for rank in rankings if rank <= cutoff)
    return hits / total_shoemarks


def cmp_all(rankings: list[int], total_shoeprints: int, total_shoemarks: int) -> None:
    """Print the S1/S5/S10/S15/S20 line in the reference's format (``parse_results.py:29-35``)."""
    values = [cmp(rankings, p, total_shoeprints, total_shoemarks) * 100 for p in S_PERCENTAGES]
    print(" ".join(f"S{p}:{v:.2f}" for p, v in zip(S_PERCENTAGES, values)))
