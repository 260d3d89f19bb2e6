#!/usr/bin/env python3
"""Kernel-time split of a config-0-scale ragged compare (300 probes of individual shapes x 1,175 gallery prints, 25 variants)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import __graft_entry__ as ge
ge.build()
from torch.profiler import profile, ProfilerActivity
from src.shoeprint_image_retrieval import engine, synth
q, g = int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1175
gallery = synth.make_gallery(1, g, 80, 59, 21)
probes, pairs = synth.make_probes(2, gallery, q, min_frac=0.4)
rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
engine.compare(probes[:8], gallery, pairs[:8], rot, scl)
torch.cuda.synchronize(); t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ranks, scores, _ = engine.compare(probes, gallery, pairs, rot, scl)
    torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{dt:.2f} s wall for {q}x{g} pairs x 25 variants -> {q*g/dt:.0f} pairs/s")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=50))
