"""Seeded FID-300-shaped dataset directories for the loader parity vectors (used by make_golden.py, which runs the
reference's Dataloader on them, and by tests/test_host_logic.py, which runs this repository's)."""

from __future__ import annotations

from pathlib import Path

import numpy as np
from PIL import Image


def write_dataset(root: Path, seed: int, gallery_sizes: list[tuple[int, int]], query_sizes: list[tuple[int, int]]) -> None:
    """``Gallery/%05d.png`` of the given (height, width) sizes, ``Query/%05d.png`` crops of random gallery prints,
    ``label_table.csv`` rows ``query_id,gallery_id`` (1-based), all from ``numpy.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    (root / "Gallery").mkdir(parents=True)
    (root / "Query").mkdir()
    prints = []
    for i, (h, w) in enumerate(gallery_sizes, start=1):
        base = rng.integers(0, 256, size=(h // 8 + 1, w // 8 + 1)).astype(np.float32)
        img = np.kron(base, np.ones((8, 8), np.float32))[:h, :w]
        img = np.clip(img * 0.6 + rng.normal(60, 25, img.shape), 0, 255).astype(np.uint8)
        prints.append(img)
        Image.fromarray(img).save(root / "Gallery" / f"{i:05d}.png")
    rows = []
    for q, (h, w) in enumerate(query_sizes, start=1):
        g = int(rng.integers(0, len(prints)))
        src = prints[g]
        h, w = min(h, src.shape[0]), min(w, src.shape[1])
        y0, x0 = int(rng.integers(0, src.shape[0] - h + 1)), int(rng.integers(0, src.shape[1] - w + 1))
        crop = np.clip(src[y0 : y0 + h, x0 : x0 + w].astype(np.float32) + rng.normal(0, 10, (h, w)), 0, 255).astype(np.uint8)
        Image.fromarray(crop).save(root / "Query" / f"{q:05d}.png")
        rows.append(f"{q},{g + 1}")
    (root / "label_table.csv").write_text("\n".join(rows) + "\n")


# (name, seed, gallery sizes, query sizes, n_clusters): one size population, and two well separated ones (any KMeans
# seed finds the same two groups; the reference's is unseeded, dataloader.py:284)
CASES = [
    ("one", 11, [(586, 270)] * 8, [(int(h), int(w)) for h, w in zip(np.linspace(300, 560, 12), np.linspace(180, 260, 12))], 1),
    ("two", 12, [(900, 420)] * 8, [(260 + 6 * i, 200 + 2 * i) for i in range(6)] + [(820 + 5 * i, 380 + 3 * i) for i in range(6)], 2),
]


def config_for(root: Path, n_clusters: int) -> dict:
    return {
        "dataset": {"dir": str(root) + "/", "type": "FID-300", "crop": [0.1, 0.2], "n_processes": 4, "n_clusters": n_clusters,
                    "cluster_minimise_tolerance": 0.05},
        "model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8], "start_block": 6, "end_block": 4,
                  "skip_blocks": [5], "minimum_dim": 300, "maximum_dim": 800},
        "comparison": {"n_processes": 4, "rotations": None, "scales": None},
    }
