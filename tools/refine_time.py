#!/usr/bin/env python3
"""Screen / refine split of the default precision mode (development aid): per-kernel CUDA-event times and the
refinement's work counters for one score_matrix pass."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import __graft_entry__ as ge

ge.build()
from src.shoeprint_image_retrieval import engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=176)
ap.add_argument("--h", type=int, default=50)
ap.add_argument("--w", type=int, default=19)
ap.add_argument("--g", type=int, default=150)
ap.add_argument("--q", type=int, default=1500)
ap.add_argument("--rot", type=int, default=12)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()

gal = synth.device_gallery(1, a.g, a.c, a.h, a.w)
prb, pairs = synth.device_probes(2, gal, a.q)
ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
rots = list(range(1, a.rot + 1)) or None
engine.collect_refine_stats = True
for i in range(a.reps + 1):
    engine.kernel_events, engine.refine_events = [], []
    engine.refine_stats(reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s = engine.score_matrix(ps, gs, rots, None, "fp16_refine")
    e1.record()
    torch.cuda.synchronize()
    screen = sum(x.elapsed_time(y) for x, y, _ in engine.kernel_events)
    refine = sum(x.elapsed_time(y) for x, y in engine.refine_events)
    flops = sum(f for _, _, f in engine.kernel_events)
    st = engine.refine_stats()
    print(f"rep {i}: total {e0.elapsed_time(e1):.1f} ms  screen {screen:.1f} ms ({flops / screen / 1e9:.0f} TFLOP/s)  refine {refine:.1f} ms  "
          f"positions {st['positions']} ({st['positions'] / (a.q * a.g):.2f}/pair)  dense records {st['dense_records']}  tiles {st['tiles']}", flush=True)
