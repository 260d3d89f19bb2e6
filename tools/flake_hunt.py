#!/usr/bin/env python3
"""Repeat one small compare() many times per precision mode and report runs whose scores differ from the first run."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine as eng, synth

c, h, w = 80, 59, 21
gallery = synth.make_gallery(101, 4, c, h, w)
probes, pairs = synth.make_probes(102, gallery, 3, min_frac=0.85)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for precision in ("fp16_refine", "fp16_fp8c", "fp16x3"):
    for resident in (False,):
        if resident:
            p = torch.from_numpy(np.stack(probes)).cuda(); g = torch.from_numpy(np.stack(gallery)).cuda()
        first, bad = None, 0
        for it in range(n):
            if resident:
                scores = eng.score_matrix(eng.MapSet.from_device(p), eng.MapSet.from_device(g), [-5], None, precision=precision)
                s = scores.cpu().numpy()[:, :4]
            else:
                _, scores, _ = eng.compare(probes, gallery, pairs, [-5], None, precision=precision)
                s = scores.cpu().numpy()
            if first is None:
                first = s
            elif not np.array_equal(first, s):
                bad += 1
                if bad <= 3:
                    d = np.argwhere(first != s)
                    print(precision, "resident" if resident else "host", "iter", it, "differs at", d.tolist(), [(float(first[tuple(i)]), float(s[tuple(i)])) for i in d], flush=True)
        print(precision, "resident" if resident else "host", "runs", n, "differing", bad, flush=True)
