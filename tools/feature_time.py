#!/usr/bin/env python3
"""Feature-stage timing helper: one batched forward of EfficientNetV2-M[:block] on synthetic 800x300 prints."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import network
block = int(sys.argv[1]) if len(sys.argv) > 1 else 6
b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = {"model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
model = network.Model(cfg, block, random_init_seed=0)
batch = np.random.default_rng(0).integers(0, 256, size=(b, 800, 300), dtype=np.uint8)
d = torch.from_numpy(batch).cuda()
for _ in range(2):  # warm-up
    model._forward_device(d, apply_clahe=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); model._forward_device(d, apply_clahe=True); model._forward_device(d, apply_clahe=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"block {block} batch {b}: {ms:.1f} ms -> {b / ms * 1e3:.0f} images/s (device-resident input)")
