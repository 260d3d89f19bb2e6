#!/usr/bin/env python3
"""BASELINE configs[3]-style run on one GPU: Q probes x G on-device gallery maps of 80x59x21, no variants,
sharded rank + top-64.  Reports pairs/s and peak device memory."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, sharding, synth
q, g = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 50000
gal = synth.device_gallery(1, g, 80, 59, 21)
prb, pairs = synth.device_probes(2, gal, q)
ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats(); t0 = time.perf_counter()
ranks, tv, ti, scores = sharding.compare_sharded(ps, gs, pairs, 0, None, None, k=64)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{q} x {g} pairs: {dt:.2f} s -> {q*g/dt/1e6:.2f} M pairs/s; rank-1 share {(ranks==1).float().mean().item():.3f}; "
      f"peak device memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB (gallery input {gal.numel()*4/2**30:.1f} GiB)")
