"""The UNMODIFIED reference as the CPU arm of the benchmark (test / measurement infrastructure only).

``vendor()`` copies the reference package from ``/root/reference`` (present in the build container only) into
the git-ignored ``baseline/_ref/`` -- which travels to the GPU box with the snapshot -- and ``time_workers()``
drives the reference's own ``_comparison_worker`` (``similarity.py:287-375``) in forked processes, the way
``compare_maps`` does (``similarity.py:146-197``), without its 1-second progress poll (``:204-212``) and without
the hang when rotations and scales are both set (``:200-204``).  Nothing of this is on the
product path; ``bench.py`` falls back to the oracle port when ``baseline/_ref`` is absent.
"""

from __future__ import annotations

import importlib.util
import multiprocessing as mp
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF_SRC = Path("/root/reference/src/shoeprint_image_retrieval")
REF_DST = ROOT / "baseline" / "_ref" / "src" / "shoeprint_image_retrieval"
_PKG = "_sir_reference_pkg"


def vendor() -> bool:
    """Copy the reference's package sources next to the repo (git-ignored); True when a copy is in place."""
    if REF_SRC.is_dir():
        REF_DST.mkdir(parents=True, exist_ok=True)
        for f in REF_SRC.glob("*.py"):
            dst = REF_DST / f.name
            if not dst.exists() or dst.read_bytes() != f.read_bytes():
                shutil.copyfile(f, dst)
    return available()


def available() -> bool:
    return (REF_DST / "similarity.py").exists()


def load_similarity():
    """The reference's ``similarity`` module, imported from ``baseline/_ref`` under a private package name."""
    if f"{_PKG}.similarity" in sys.modules:
        return sys.modules[f"{_PKG}.similarity"]
    if not available():
        raise ImportError("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    spec = importlib.util.spec_from_file_location(_PKG, REF_DST / "__init__.py", submodule_search_locations=[str(REF_DST)])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[_PKG] = pkg
    spec.loader.exec_module(pkg)
    return importlib.import_module(f"{_PKG}.similarity")


def time_workers(probe_maps, gallery_maps, matching_pairs, rotations, scales, n_procs: int | None = None):
    """Run the reference's workers over ``n_procs`` contiguous probe chunks; returns (ranks int32 [Q], seconds, n_procs)."""
    sim = load_similarity()
    q = len(probe_maps)
    n_procs = max(1, min(n_procs or os.cpu_count() or 1, q))
    ctx = mp.get_context("fork")
    shared = []
    for g in gallery_maps:  # similarity.py:164-176: every gallery map in shared memory, re-viewed by the workers
        arr = mp.Array("f", int(g.size))
        np.frombuffer(arr.get_obj(), dtype=np.float32)[:] = np.ascontiguousarray(g, dtype=np.float32).ravel()
        shared.append((arr, g.shape))
    trackers = sim.MultiProcessingTrackers(q)
    base, extra = divmod(q, n_procs)
    procs, start = [], 0
    t0 = time.perf_counter()
    for i in range(n_procs):
        end = start + base + (1 if i < extra else 0)
        pr = ctx.Process(target=sim._comparison_worker,
                         args=(list(probe_maps[start:end]), shared, (start, end), list(matching_pairs), trackers, rotations, scales))
        pr.start()
        procs.append(pr)
        start = end
    import queue as _queue

    drained = 0
    while drained < q:  # the workers' queue must be drained or they block at exit
        try:
            trackers.queue.get(timeout=1.0)
            drained += 1
        except _queue.Empty:
            if not any(pr.is_alive() for pr in procs) and trackers.queue.empty():
                raise RuntimeError("a reference worker died before reporting its ranks") from None
    for pr in procs:
        pr.join()
    dt = time.perf_counter() - t0
    return np.frombuffer(trackers.rankings.get_obj(), dtype=np.int32).copy(), dt, n_procs
