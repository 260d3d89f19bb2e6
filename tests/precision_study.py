#!/usr/bin/env python3
"""Score error of the tensor-core precision modes against the fp32 CUDA-core evaluation (and, for a
subsample, the float64 CPU oracle) at the reference map shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, synth
from oracle import compare as ocmp
for c, h, w, g, q in [(176, 50, 19, 24, 16), (80, 59, 21, 24, 16)]:
    gal = synth.device_gallery(3, g, c, h, w); prb, pairs = synth.device_probes(4, gal, q)
    ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
    rot = [-10, 10]
    ref32 = engine.score_matrix(ps, gs, rot, None, "fp32_simt").cpu().numpy()
    sub_q, sub_g = 3, 4
    _, want = ocmp.compare_maps_oracle([m for m in prb[:sub_q].cpu().numpy()], [m for m in gal[:sub_g].cpu().numpy()], [0] * sub_q, rot, None)
    for mode in ("fp16_refine", "fp16x3", "fp16_fp8c", "fp16x1"):
        s = engine.score_matrix(ps, gs, rot, None, mode).cpu().numpy()
        e32 = np.abs(s - ref32) / np.maximum(np.abs(ref32), 1e-3)
        e64 = np.abs(s[:sub_q, :sub_g] - want) / np.maximum(np.abs(want), 1e-3)
        print(f"C={c} {h}x{w} {mode:10s}: vs fp32 SIMT max {e32.max():.2e} mean {e32.mean():.2e} | vs float64 oracle ({sub_q}x{sub_g}) max {e64.max():.2e}", flush=True)
    e = np.abs(ref32[:sub_q, :sub_g] - want) / np.maximum(np.abs(want), 1e-3)
    print(f"C={c} {h}x{w} fp32_simt : vs float64 oracle max {e.max():.2e}; score range [{ref32.min():.3f}, {ref32.max():.3f}]", flush=True)
