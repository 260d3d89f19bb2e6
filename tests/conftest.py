"""Shared pytest configuration: markers, repo root on sys.path, golden-vector fixture."""

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    with np.load(ROOT / "tests" / "golden" / "reference_vectors.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_edge():
    """Edge-case vectors from the unmodified reference: dead / constant channels, ReLU-like flat regions, anti-correlated
    probes (tests/golden/make_golden.py, edge_cases)."""
    with np.load(ROOT / "tests" / "golden" / "reference_vectors_edge.npz") as z:
        return {k: z[k] for k in z.files}
