#!/usr/bin/env python3
"""Where a ragged (configs[0]-shaped) compare pass spends its time: host profile (cProfile) and summed CUDA kernel times
(torch.profiler / CUPTI).  Development aid."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import __graft_entry__ as ge

ge.build()
from src.shoeprint_image_retrieval import engine, synth

q, g = int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1175
gallery = synth.make_gallery(1, g, 80, 59, 21)
probes, pairs = synth.make_probes(2, gallery, q, min_frac=0.4)
rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
engine.compare(probes, gallery, pairs, rot, scl)
torch.cuda.synchronize()
n0 = engine.launch_counter.n
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
engine.compare(probes, gallery, pairs, rot, scl)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
pr.disable()
print(f"wall {time.perf_counter() - t0:.3f} s, host returned after {t_host:.3f} s, libsir launches {engine.launch_counter.n - n0}")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    engine.compare(probes, gallery, pairs, rot, scl)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
