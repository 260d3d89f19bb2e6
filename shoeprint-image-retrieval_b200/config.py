"""``run.toml`` schema and loader (reference: ``config.py:11-64``).

Same keys and the same normalisation as the reference: an empty string for
``comparison.rotations`` / ``comparison.scales`` means "no such variants" and is turned into
``None`` (``config.py:60-63``).  Extra, optional keys understood by this implementation
(absent -> reference behaviour):

* ``comparison.precision``  ``"fp16_refine"`` (default, parity grade) | ``"fp16_fp8c"`` | ``"fp16x3"`` | ``"fp16x1"`` | ``"fp32_simt"``
* ``comparison.top_k``      length of the per-probe candidate list kept next to the ranks
"""

from __future__ import annotations

from pathlib import Path
from typing import TypedDict

import toml

from .customtypes import DatasetTypeType


class DatasetConfig(TypedDict):
    dir: str
    type: DatasetTypeType
    crop: list[float]
    n_processes: int
    n_clusters: int
    cluster_minimise_tolerance: float


class ModelConfig(TypedDict):
    type: str
    clahe_clip_limit: float
    clahe_tile_grid_size: list[int]
    start_block: int
    end_block: int
    skip_blocks: list[int]
    minimum_dim: int
    maximum_dim: int


class ComparisonConfig(TypedDict, total=False):
    n_processes: int
    rotations: list[int] | None
    scales: list[float] | None
    precision: str
    top_k: int


class Config(TypedDict):
    dataset: DatasetConfig
    model: ModelConfig
    comparison: ComparisonConfig


def load_config(config_file: Path | str) -> Config:
    """Parse ``config_file`` (TOML) into a :class:`Config`."""
    parsed = toml.loads(Path(config_file).read_text())
    comparison = parsed["comparison"]
    for key in ("rotations", "scales"):
        if comparison.get(key) == "":
            comparison[key] = None
    return Config(parsed)  # type: ignore[typeddict-item]
