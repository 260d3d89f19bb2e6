#!/usr/bin/env python3
"""Ad-hoc timing of a config-1-like ragged workload (every probe its own shape, 25 variants)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, synth
q, g = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 256
gallery = synth.make_gallery(1, g, 80, 59, 21)
probes, pairs = synth.make_probes(2, gallery, q, min_frac=0.4)
rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ranks, scores, _ = engine.compare(probes, gallery, pairs, rot, scl)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt:.2f} s for {q}x{g} pairs x 25 variants -> {q*g/dt:.0f} pairs/s; rank-1 share {(ranks==1).mean():.2f}", flush=True)
