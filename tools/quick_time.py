#!/usr/bin/env python3
"""Ad-hoc device timing of the correlation kernel (development aid, not the bench contract)."""
import argparse
import sys
import time
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
ap.add_argument("--q", type=int, default=256)
ap.add_argument("--rot", type=int, default=0)
ap.add_argument("--prec", default="fp16x3")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

gal = synth.device_gallery(1, a.g, a.c, a.h, a.w)
prb, pairs = synth.device_probes(2, gal, a.q)
ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
rots = list(range(1, a.rot + 1)) or None
for i in range(a.reps + 1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s = engine.score_matrix(ps, gs, rots, None, a.prec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    v = 1 + a.rot
    flops = 2.0 * a.c * ((a.h - 4) * (a.w - 4)) ** 2 * a.q * a.g * v
    print(f"rep {i}: {ms:.2f} ms  pairs/s {a.q * a.g / ms * 1e3:.0f}  algorithmic {flops / ms / 1e9:.1f} TFLOP/s ({a.prec})", flush=True)
acc = (s.argmax(1) == pairs.long()).float().mean().item()
print("top-1 accuracy", acc)
