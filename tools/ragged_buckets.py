import sys, time
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import engine, synth
gallery = synth.make_gallery(1, 1175, 80, 59, 21)
probes, pairs = synth.make_probes(2, gallery, 300, min_frac=0.4)
rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
orig = engine._score_one_bucket
log = []
def spy(members, gops, g0, scores, mode, flip, dev, approx=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig(members, gops, g0, scores, mode, flip, dev, approx)
    torch.cuda.synchronize()
    hb = max(hw[0] for hw, _ in members); wb = max(hw[1] for hw, _ in members)
    log.append((len(members), sum(b.ncols for _, b in members), flip, hb, wb, time.perf_counter() - t0))
engine._score_one_bucket = spy
engine.compare(probes, gallery, pairs, rot, scl)
engine.compare(probes, gallery, pairs, rot, scl)
half = len(log) // 2
for row in log[half:]:
    print("shapes %3d cols %4d flip %s bucket %dx%d  %.1f ms" % (row[0], row[1], row[2], row[3], row[4], row[5] * 1e3))
print("sum", sum(r[5] for r in log[half:]))
