#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Usage (build container only; ``/root/reference`` does not exist on the GPU box):
    python tests/golden/make_golden.py

Imports ``/root/reference/src/shoeprint_image_retrieval/{similarity,parse_results}.py`` and
records inputs + outputs of ``normxcorr``, ``get_similarity``, ``_apply_transformations``,
``_comparison_worker`` (called in-process: ``compare_maps`` itself never returns when both
rotations and scales are set) and ``cmp`` on seeded synthetic arrays.
The committed ``*.npz`` files are what ``tests/`` check the oracle and the CUDA path against.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

sys.path.insert(0, str(REF))
from src.shoeprint_image_retrieval import parse_results as ref_parse  # noqa: E402
from src.shoeprint_image_retrieval import similarity as ref_sim  # noqa: E402


def smooth_field(rng, c, h, w, passes=2):
    """Print-like feature maps: low-pass noise + texture, no exactly flat regions."""
    a = rng.standard_normal((c, h + 4, w + 4)).astype(np.float32)
    for _ in range(passes):
        a = (a + np.roll(a, 1, 1) + np.roll(a, 1, 2) + np.roll(a, -1, 1) + np.roll(a, -1, 2)) / 5
    a = a[:, 2:-2, 2:-2] * 6 + 0.15 * rng.standard_normal((c, h, w)).astype(np.float32)
    return np.ascontiguousarray(a.astype(np.float32))


def make_probe(rng, gal, hq, wq, noise=0.3):
    c, h, w = gal.shape
    y0 = int(rng.integers(0, h - hq + 1))
    x0 = int(rng.integers(0, w - wq + 1))
    crop = gal[:, y0 : y0 + hq, x0 : x0 + wq]
    return np.ascontiguousarray((crop + noise * rng.standard_normal(crop.shape)).astype(np.float32))


class _FakeShared:
    """Stands in for multiprocessing.Array: the worker only calls ``.get_obj()`` (similarity.py:312-317)."""

    def __init__(self, arr):
        self._a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)

    def get_obj(self):
        return self._a


def run_worker(probes, gallery, pairs, rotations, scales):
    trackers = ref_sim.MultiProcessingTrackers(len(probes))
    shared = [(_FakeShared(g), g.shape) for g in gallery]
    ref_sim._comparison_worker(list(probes), shared, (0, len(probes)), pairs, trackers, rotations, scales)
    ranks = np.frombuffer(trackers.rankings.get_obj(), dtype=np.int32).copy()
    counter = trackers.counter.value
    while not trackers.queue.empty():
        trackers.queue.get()
    return ranks, counter


def ref_scores(probes, gallery, rotations, scales):
    """Score matrix built with the reference's own functions, mirroring similarity.py:321-367."""
    if rotations is not None and scales is not None:
        rot = ref_sim._apply_transformations([list(probes)], rotations, list(probes), "rotate")
        lists = ref_sim._apply_transformations(rot, scales, list(probes), "scale")
    elif rotations is not None:
        lists = ref_sim._apply_transformations([list(probes)], rotations, list(probes), "rotate")
    elif scales is not None:
        lists = ref_sim._apply_transformations([list(probes)], scales, list(probes), "scale")
    else:
        lists = [list(probes)]
    best = np.zeros((len(probes), len(gallery)), dtype=np.float32)
    for marks in lists:
        for qi, mark in enumerate(marks):
            for gi, prnt in enumerate(gallery):
                sim = ref_sim.get_similarity(mark, prnt)
                if sim > best[qi, gi]:
                    best[qi, gi] = sim
    return best, len(lists)


def main() -> None:
    rng = np.random.default_rng(20261018)
    out: dict[str, np.ndarray] = {}

    # 1. normxcorr: (template, image) size pairs incl. even/odd, template larger than image
    sizes = [((7, 5), (11, 9)), ((6, 4), (11, 9)), ((13, 5), (11, 9)), ((12, 12), (8, 8)),
             ((9, 9), (9, 9)), ((15, 7), (15, 7)), ((3, 2), (10, 6))]
    for i, (ts, gs) in enumerate(sizes):
        t = rng.standard_normal(ts).astype(np.float32) * 3 + 1
        g = rng.standard_normal(gs).astype(np.float32) * 2 - 0.5
        out[f"nx{i}_t"], out[f"nx{i}_g"] = t, g
        out[f"nx{i}_out"] = np.asarray(ref_sim.normxcorr(t, g, "same"), dtype=np.float64)
    out["nx_count"] = np.array(len(sizes))

    # 2. get_similarity on small multi-channel maps
    gs_cases = [((6, 14, 11), (6, 14, 11)), ((5, 10, 9), (5, 16, 12)), ((4, 17, 13), (4, 13, 10)), ((8, 12, 12), (8, 12, 12))]
    for i, (ps, gshape) in enumerate(gs_cases):
        gal = smooth_field(rng, *gshape)
        if ps[1] <= gshape[1] and ps[2] <= gshape[2]:
            prb = make_probe(rng, gal, ps[1], ps[2])
        else:
            prb = smooth_field(rng, *ps)
        out[f"gs{i}_p"], out[f"gs{i}_g"] = prb, gal
        out[f"gs{i}_out"] = np.array(float(ref_sim.get_similarity(prb, gal)))
    out["gs_count"] = np.array(len(gs_cases))

    # 3. variants through _apply_transformations
    rot_angles = [-30, -15, -9, -3, 3, 9, 15, 25, 90, 180, 270, 360]
    scl = [1.02, 1.04, 1.08, 0.9, 0.5, 1.3, 2.1]
    shapes = [(3, 13, 9), (2, 12, 12), (3, 59, 21), (2, 50, 19), (2, 8, 30)]
    for i, shp in enumerate(shapes):
        m = smooth_field(rng, *shp)
        out[f"var{i}_in"] = m
        lists = ref_sim._apply_transformations([[m]], rot_angles, [m], "rotate")
        for j, ang in enumerate(rot_angles):
            out[f"var{i}_rot{j}"] = lists[1 + j][0]
        lists = ref_sim._apply_transformations([[m]], scl, [m], "scale")
        for j, s in enumerate(scl):
            out[f"var{i}_scl{j}"] = lists[1 + j][0]
    out["var_count"] = np.array(len(shapes))
    out["var_rot_angles"] = np.array(rot_angles, dtype=np.float64)
    out["var_scales"] = np.array(scl, dtype=np.float64)

    # 4. whole compare pass, four variant modes, ragged probes
    c, hg, wg, q, g = 4, 16, 12, 6, 9
    gallery = [smooth_field(rng, c, hg, wg) for _ in range(g)]
    pairs = [int(x) for x in rng.integers(0, g, size=q)]
    probes = []
    for qi in range(q):
        hq = int(rng.integers(9, hg + 1))
        wq = int(rng.integers(8, wg + 1))
        probes.append(make_probe(rng, gallery[pairs[qi]], hq, wq))
    out["cmp_gallery"] = np.stack(gallery)
    out["cmp_pairs"] = np.array(pairs, dtype=np.int64)
    for qi, p in enumerate(probes):
        out[f"cmp_probe{qi}"] = p
    out["cmp_q"] = np.array(q)
    modes = {"none": (None, None), "rot": ([-15, 9, 180], None), "scl": (None, [1.04, 1.2]),
             "both": ([-9, 15], [1.04, 1.2])}
    for name, (rots, scs) in modes.items():
        ranks, counter = run_worker(probes, gallery, pairs, rots, scs)
        scores, nlists = ref_scores(probes, gallery, rots, scs)
        out[f"cmp_{name}_ranks"] = ranks
        out[f"cmp_{name}_scores"] = scores
        out[f"cmp_{name}_counter"] = np.array(counter)
        out[f"cmp_{name}_nvariants"] = np.array(nlists)
        out[f"cmp_{name}_rot"] = np.array([] if rots is None else rots, dtype=np.float64)
        out[f"cmp_{name}_scl"] = np.array([] if scs is None else scs, dtype=np.float64)

    # 5. S-scores
    ranks = rng.integers(1, 400, size=57).astype(np.int32)
    out["s_ranks"] = ranks
    out["s_total_prints"], out["s_total_marks"] = np.array(1175), np.array(300)
    out["s_values"] = np.array([ref_parse.cmp(list(ranks), p, 1175, 300) for p in (1, 5, 10, 15, 20)])

    np.savez_compressed(OUT / "reference_vectors.npz", **out)
    print("wrote", OUT / "reference_vectors.npz", sum(v.nbytes for v in out.values()), "bytes raw")
    edge_cases()
    loader_cases()


def loader_cases() -> None:
    """Third file: what the reference's Dataloader (dataloader.py:29-469) hands to the hot path for two generated
    FID-300-shaped directories -- per cluster its scale, backbone block, file list, matching pairs and the loaded
    (cropped, LANCZOS-resized) uint8 images."""
    import hashlib
    import tempfile

    from src.shoeprint_image_retrieval.dataloader import Dataloader as RefLoader

    sys.path.insert(0, str(OUT))
    import synth_dataset

    out: dict[str, np.ndarray] = {}
    for name, seed, gsizes, qsizes, n_clusters in synth_dataset.CASES:
        with tempfile.TemporaryDirectory() as tmp:
            root = Path(tmp) / name
            synth_dataset.write_dataset(root, seed, gsizes, qsizes)
            loader = RefLoader(synth_dataset.config_for(root, n_clusters))
            out[f"ld_{name}_nclusters"] = np.array(loader.num_clusters)
            out[f"ld_{name}_scales"] = np.array(loader.scales, dtype=np.float64)
            out[f"ld_{name}_blocks"] = np.array(loader.blocks, dtype=np.int64)
            for k, (marks, prints, pairs, block) in enumerate(loader):
                out[f"ld_{name}_c{k}_files"] = np.array(sorted(loader.clusters[k]))
                out[f"ld_{name}_c{k}_pairs"] = np.array(pairs, dtype=np.int64)
                out[f"ld_{name}_c{k}_block"] = np.array(block)
                out[f"ld_{name}_c{k}_mark0"], out[f"ld_{name}_c{k}_print0"] = marks[0], prints[0]
                digest = hashlib.sha256()
                for im in list(marks) + list(prints):
                    digest.update(repr(im.shape).encode() + np.ascontiguousarray(im).tobytes())
                out[f"ld_{name}_c{k}_sha256"] = np.array(digest.hexdigest())
                out[f"ld_{name}_c{k}_counts"] = np.array([len(marks), len(prints)])
    np.savez_compressed(OUT / "reference_loader.npz", **out)
    print("wrote", OUT / "reference_loader.npz", {k: v.tolist() for k, v in out.items() if k.endswith(("scales", "blocks", "nclusters"))})


def edge_cases() -> None:
    """Second file (own seed, so reference_vectors.npz stays bit-identical): inputs real post-activation feature maps
    produce and synthetic smooth fields avoid -- dead (all-zero) and constant channels on either side, ReLU-like maps
    with flat halves, an anti-correlated probe -- through get_similarity and through the worker with rotations."""
    rng = np.random.default_rng(20261019)
    out: dict[str, np.ndarray] = {}
    g = smooth_field(rng, 5, 16, 12)
    p = np.ascontiguousarray(g + 0.3 * rng.standard_normal(g.shape).astype(np.float32))
    cases = []
    g_dead = g.copy(); g_dead[1] = 0.0
    cases.append((p, g_dead))                       # dead gallery channel
    p_dead = p.copy(); p_dead[2] = 0.0
    cases.append((p_dead, g))                       # dead probe channel
    cases.append((p_dead, g_dead))                  # both
    g_const = g.copy(); g_const[1] = 0.7321
    cases.append((p, g_const))                      # constant non-zero gallery channel
    g_relu = np.maximum(g, 0); g_relu[3, :, :6] = 0
    cases.append((np.maximum(p, 0), np.ascontiguousarray(g_relu)))  # ReLU-like, one channel flat over half the map
    cases.append((np.ascontiguousarray(-p), g))     # anti-correlated probe
    small = make_probe(rng, g, 9, 8)
    cases.append((np.ascontiguousarray(-small), g_dead))
    for i, (a, b) in enumerate(cases):
        out[f"eg{i}_p"], out[f"eg{i}_g"] = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
        out[f"eg{i}_out"] = np.array(float(ref_sim.get_similarity(out[f"eg{i}_p"], out[f"eg{i}_g"])))
    out["eg_count"] = np.array(len(cases))
    # compare pass over edge maps
    gallery = [smooth_field(rng, 4, 15, 12) for _ in range(6)]
    gallery[1][2] = 0.0
    gallery[3] = np.ascontiguousarray(np.maximum(gallery[3], 0))
    gallery[4][0] = 1.5
    pairs = [0, 1, 3, 4, 5]
    probes = [make_probe(rng, gallery[t], int(rng.integers(9, 16)), int(rng.integers(8, 13))) for t in pairs]
    probes[1][2] = 0.0
    probes[4] = np.ascontiguousarray(-probes[4])    # its true match is anti-correlated: ranks low, scores near the 0 floor
    out["ecmp_gallery"] = np.stack(gallery)
    out["ecmp_pairs"] = np.array(pairs, dtype=np.int64)
    for qi, pr in enumerate(probes):
        out[f"ecmp_probe{qi}"] = pr
    out["ecmp_q"] = np.array(len(probes))
    rots = [-12, 8]
    ranks, _ = run_worker(probes, gallery, pairs, rots, None)
    scores, _ = ref_scores(probes, gallery, rots, None)
    out["ecmp_rot"] = np.array(rots, dtype=np.float64)
    out["ecmp_ranks"], out["ecmp_scores"] = ranks, scores
    np.savez_compressed(OUT / "reference_vectors_edge.npz", **out)
    print("wrote", OUT / "reference_vectors_edge.npz", [float(out[f"eg{i}_out"]) for i in range(len(cases))], ranks)


if __name__ == "__main__":
    main()
