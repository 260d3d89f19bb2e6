#!/usr/bin/env python3
"""Host profile of one similarity.compare_maps call on the headline workload (pageable numpy lists)."""
import cProfile, contextlib, io, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, similarity, synth
q, g, c, h, w = 1500, 150, 176, 50, 19
gallery = synth.make_gallery(1, g, c, h, w)
probes, pairs = synth.make_probes(2, gallery, q, min_frac=1.0) if hasattr(synth, "make_probes") else None
probes = [np.ascontiguousarray(p) for p in probes]
cfg = {"comparison": {"n_processes": 1, "rotations": list(range(-30, 31, 5)), "scales": None}}
cfg["comparison"]["rotations"].remove(0)
def run():
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        return similarity.compare_maps(probes, gallery, pairs, cfg)
for _ in range(2):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("e2e wall", [round(t * 1e3, 1) for t in ts])
ps, gs = engine.MapSet.from_host(probes), engine.MapSet.from_host(gallery)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); engine.compare(ps, gs, pairs, cfg["comparison"]["rotations"], None); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("resident wall", [round(t * 1e3, 1) for t in ts])
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
