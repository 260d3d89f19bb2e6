#!/usr/bin/env python3
"""Where the host time of Model.get_multiple_feature_maps goes (512 synthetic 800x300 prints)."""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import network
cfg = {"model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
model = network.Model(cfg, 6, random_init_seed=0)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
imgs = [rng.integers(0, 256, size=(800, 300), dtype=np.uint8) for _ in range(n)]
model.get_multiple_feature_maps(imgs[:64], progress=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
maps = model.get_multiple_feature_maps(imgs, progress=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n} images: {dt * 1e3:.0f} ms -> {n / dt:.0f} images/s")
pr = cProfile.Profile()
pr.enable()
model.get_multiple_feature_maps(imgs[:256], progress=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
