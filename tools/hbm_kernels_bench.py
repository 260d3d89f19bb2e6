#!/usr/bin/env python3
"""Achieved HBM GB/s of the bandwidth-bound kernels (K4, K5, K6, K8) against the measured copy peak.
Algorithmic bytes per SURVEY.md 8d / DESIGN.md section 4; CUDA events on the launching stream."""
import ctypes as C, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import __graft_entry__ as ge
ge.build()
from src.shoeprint_image_retrieval import _native as nat, engine

peak = 6548.2
f = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
if f.exists():
    peak = json.loads(f.read_text()).get("hbm_gbs", peak)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.fill_(1)  # > L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

rows = []
def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    rows.append({"kernel": name, "algorithmic_bytes": nbytes, "ms": ms, "GB/s": gbs, "frac_of_measured_peak": gbs / peak})
    print(f"{name:28s} {nbytes/1e6:10.1f} MB {ms:8.3f} ms {gbs:8.0f} GB/s  {100*gbs/peak:5.1f} % of {peak:.0f}", flush=True)

# K8 rank / top-k at gallery-scaling size: Q=1000 x G=100,000 scores
q, g, k = 1000, 100000, 64
scores = torch.rand((q, g), device="cuda")
true = torch.randint(0, g, (q,), device="cuda", dtype=torch.int32)
report("rank_topk (K8) k=64", 4 * q * g + q * (8 * k + 8), timed(lambda: engine.rank_true_matches(scores, true, k)))
report("rank_topk (K8) k=0", 4 * q * g + q * 8, timed(lambda: engine.rank_true_matches(scores, true, 0)))

# K5g gallery pack, K6 window norm, K4 rotate, K5t template pack at 8,192 maps of 80x59x21
n, c, h, w = 8192, 80, 59, 21
maps = torch.randn((n, c, h, w), device="cuda")
grp = engine.MapGroup(maps, torch.arange(n))
ops = engine.GalleryOperands.pack(grp, keep_fp32=False)
hp, wp = h - 4, w - 4
wpitch = ops.ghi.shape[3]
report("gallery_pack (K5g)", n * c * (4 * h * w + 4 * hp * wpitch), timed(lambda: engine.GalleryOperands.pack(grp, keep_fp32=False)))
report("gallery_pack_f32 (K5g, default mode)", n * c * (4 * h * w + 8 * hp * wpitch),
       timed(lambda: engine.GalleryOperands.pack(grp, keep_fp32=False, with_f32=True)))
def rn():
    ops._rnorm.clear(); ops.rnorm(hp, wp, False)
report("window_rnorm (K6)", n * c * (4 * hp * wpitch + 4 * hp * wp), timed(rn))
report("rotate (K4r)", n * c * 8 * h * w, timed(lambda: engine.make_variant(maps, 7.0, None)))
report("resize 1.08 (K4s, 2 passes)", n * c * 4 * (h * w + 2 * h * 22 + 63 * 22), timed(lambda: engine.make_variant(maps, None, 1.08)))
kpad = nat.lib.sir_template_kpad_fp8c(hp, wp)
thi = torch.empty((c, n, kpad), dtype=torch.float16, device="cuda"); t8b = torch.empty((c, n, kpad), dtype=torch.uint8, device="cuda"); t8l = torch.empty_like(t8b)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
report("template_pack fp8c (K5t)", n * c * (4 * h * w + 4 * kpad),
       timed(lambda: nat.check(nat.lib.sir_template_pack_fp8c(C.c_void_p(maps.data_ptr()), n, c, h, w, 0, n, C.c_void_p(thi.data_ptr()), C.c_void_p(t8b.data_ptr()), C.c_void_p(t8l.data_ptr()), st))))
for rot in (None, 7.0):
    gm = engine._gather_map(maps.device, h, w, rot, False)
    from src.shoeprint_image_retrieval.engine import _ptr
    kp = int(nat.lib.sir_template_kpad(hp, wp))
    th = torch.empty((c, n, kp), dtype=torch.float16, device="cuda"); t32p = torch.empty((c, n, kp), dtype=torch.float32, device="cuda")
    report(f"template_pack_screen rot={rot} (K5t, default mode)", n * c * (4 * h * w + 6 * kp),
           timed(lambda: nat.check(nat.lib.sir_template_pack_screen(C.c_void_p(maps.data_ptr()), n, c, h, w, hp, wp, 0, n, C.c_void_p(th.data_ptr()),
                                                                    C.c_void_p(t32p.data_ptr()), _ptr(gm), st))))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/r02_hbm_kernels.json").write_text(json.dumps(rows, indent=1))
