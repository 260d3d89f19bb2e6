import sys, time, io, contextlib
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, similarity, synth
q, g, c, h, w = 300, 1175, 80, 59, 21
rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
gallery = synth.make_gallery(1, g, c, h, w)
probes, pairs = synth.make_probes(2, gallery, q, min_frac=0.4)
cfg = {"comparison": {"n_processes": 1, "rotations": rot, "scales": scl}}
base = None
for it in range(3):
    sink = io.StringIO()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        ranks = similarity.compare_maps(probes, gallery, pairs, cfg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pass", it, round(dt, 3), "s", round(q * g / dt), "pairs/s rank1", float((ranks == 1).mean()), "sum", int(ranks.sum()), flush=True)
