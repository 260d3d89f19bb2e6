#!/usr/bin/env python3
"""Per-kernel time split of one batched feature-stage forward (CUPTI through torch.profiler; not a bench number)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import network
from torch.profiler import profile, ProfilerActivity
block = int(sys.argv[1]) if len(sys.argv) > 1 else 6
b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mtype = sys.argv[3] if len(sys.argv) > 3 else "EfficientNetV2_M"
cfg = {"model": {"type": mtype, "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
model = network.Model(cfg, block, random_init_seed=0)
batch = np.random.default_rng(0).integers(0, 256, size=(b, 800, 300), dtype=np.uint8)
model._forward_uint8(batch, apply_clahe=True); torch.cuda.synchronize()
model.program.launch_log = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model._forward_uint8(batch, apply_clahe=True); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "sir::" in e.name], key=lambda e: e.time_range.start)
body = [e for e in evs if "clahe" not in e.name and "nhwc_to_nchw" not in e.name and "::pack_weights_kernel" not in e.name]
labels = model.program.launch_log
assert len(body) == len(labels), (len(body), len(labels))
import collections
agg = collections.OrderedDict()
for e, d in zip(body, labels):
    short = __import__("re").search(r"(\w+_kernel(?:<[^>]*>)?)", e.name).group(1)
    key = (short, d)
    n, t = agg.get(key, (0, 0.0))
    agg[key] = (n + 1, t + e.device_time)
print(f"{'kernel':28s} {'layer':38s} {'n':>3s} {'avg us':>9s} {'total ms':>9s}")
for (short, d), (n, t) in agg.items():
    print(f"{short:28s} {d:38s} {n:3d} {t / n:9.1f} {t / 1e3:9.3f}")
