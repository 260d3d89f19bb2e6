#!/usr/bin/env python3
"""Benchmark of the matching hot path: probe x gallery pairs scored per second.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp16_refine]

Headline workload (BASELINE.json configs[1]): synthetic WVU2019-shaped set -- Q = 1,500 probes x
G = 150 gallery prints PER GPU, feature maps of the default backbone cut at block 6 on 800x300
images (C = 176, 50 x 19), rotation sweep -30..30 step 5 degrees (12 angles + the unrotated probe =
13 variants).  A *step* is one full compare pass over that batch: gallery pack, variant
generation, template pack, window norms, correlation over all variants with the fused max, exact
re-evaluation of the maximum candidates, rank of the true match and top-k (+ the NCCL merge when N > 1).
A pair = one (probe, gallery print), all its variants included.

* ``value``  pairs/s with the float32 feature maps already resident in HBM (CUDA events, max over ranks)
* ``e2e``    the same pass through the reference-facing API ``similarity.compare_maps`` from ordinary
             (pageable) numpy feature-map lists to int32 ranks on the host, H2D and D2H inside the timed region
* ``roofline`` for the dominant kernel (``ncc_tc_kernel``, the tensor-core screening pass): algorithmic
             FLOPs / mean launch time, from CUDA events around every launch
* ``cpu_baseline`` the UNMODIFIED reference (``baseline/_ref``, its ``_comparison_worker`` in one forked
             process per host core; the oracle port when that copy is absent) on a bounded sample
* ``config2`` BASELINE configs[2]: the same images at block 6 (176x30x11 maps), 40 variants, same path as ``config0``
* ``config0`` BASELINE configs[0]: the default run.toml on a FID-300-shaped set (300 ragged probes x 1,175 gallery x 25
             variants) through ``compare_maps`` -- the multi-shape bucket path real data takes (N = 1 only)
* ``config4`` BASELINE configs[3]: 1,000 probes x 12,500 on-device gallery maps of 80x59x21 PER GPU (at N = 8
             that is the 1,000 x 100,000 target), V = 1 and V = 13, with the NCCL merge timed phase by phase
* ``precision_study`` BASELINE configs[4]: 176x68x132 maps (1024x2048 inputs), every precision mode against the
             float32 CUDA-core evaluation and a CPU-oracle subsample
* ``feature_stage`` images/s of the backbone (side measurement) next to torch + cuDNN on the same GPU

Multi-GPU: one process per GPU (torchrun), gallery sharded by contiguous index range (150 prints per GPU for
the headline: weak scaling), probes replicated; two tiny all-reduces + one all-gather per step.
"""

from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# ----------------------------------------------------------------------------- workloads
WORKLOAD = {
    "name": "configs[1]: WVU2019-shaped, EfficientNetV2-M block 6 maps 176x50x19, rotations -30..30 step 5",
    "Q": 1500,
    "G_per_gpu": 150,
    "C": 176,
    "h": 50,
    "w": 19,
    "rotations": [a for a in range(-30, 31, 5) if a != 0],
    "scales": None,
    "top_k": 20,
}
CONFIG4 = {"name": "configs[3]: gallery scaling, on-device maps 80x59x21", "Q": 1000, "G_per_gpu": 12500, "C": 80, "h": 59, "w": 21,
           "top_k": 64, "rotations13": [a for a in range(-30, 31, 5) if a != 0]}
CONFIG5 = {"name": "configs[4]: 1024x2048 inputs, block 6 maps 176x68x132, precision study", "Q": 64, "G_per_gpu": 1250, "C": 176,
           "h": 68, "w": 132, "top_k": 20}
PRECISIONS = ["fp16_refine", "fp16_fp8c", "fp16x3", "fp16x1", "fp32_simt"]
DTYPES = {
    "fp16_refine": "fp16 tensor-core screening (1 MMA per K step, fp32 accumulate) + exact float32 re-evaluation of the maximum candidates",
    "fp16x3": "fp16 hi/lo split x3 MMAs, fp32 accumulate (fp32-grade)",
    "fp16x1": "fp16, fp32 accumulate",
    "fp32_simt": "fp32",
    "fp16_fp8c": "fp16 hi*hi + fp8 e4m3 correction MMAs, fp32 accumulate (fp32-grade within 1e-4)",
}


def _config(world: int, precision: str) -> dict:
    """The ``config`` object of the JSON line -- the same for both arms."""
    w = WORKLOAD
    return {
        "workload": w["name"], "Q": w["Q"], "G": w["G_per_gpu"] * world, "G_per_gpu": w["G_per_gpu"], "C": w["C"],
        "map_hw": [w["h"], w["w"]], "variants": 1 + len(w["rotations"]), "top_k": w["top_k"], "precision": precision,
        "sharding": f"gallery x{world}" if world > 1 else "none",
        "l2": "inputs (1.0 GB of probe maps + 13 variants) exceed the 126 MB L2; no explicit flush",
    }


def _peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"bf16": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "hbm": float(d.get("hbm_gbs", 6500.0)),
                "src": "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a ~0.5 s step)"}
    return {"bf16": 1400.0, "hbm": 6500.0, "src": "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"}


def _ncu_traffic() -> tuple[float | None, str | None]:
    """DRAM bytes of one ncc_tc_kernel launch from the committed `ncu --set full` capture (profiles/)."""
    for name in ("r02_ncu_full_ncc_tc_kernel_final.txt", "r02_ncu_full_ncc_tc_kernel_screen.txt", "r01_ncu_full_ncc_tc_kernel_final.txt"):
        f = ROOT / "profiles" / name
        if not f.exists():
            continue
        total = 0.0
        for line in f.read_text().splitlines():
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if line.startswith(key + " ["):
                    unit = line.split("[")[1].split("]")[0]
                    val = float(line.split("=")[1])
                    total += val * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        return total, f"ncu --set full capture of the bench --profile launch (profiles/{name}; 3,328 columns x 150 gallery, the bench's own launch is 19,500 columns wide and moves proportionally more)"
    return None, None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        self.index = index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_baseline(budget_s: float = 18.0, procs: int | None = None) -> dict:
    """The reference's own CPU path (``baseline/_ref``: ``_comparison_worker`` forked over all host cores; the oracle
    port when the copy is absent) over a bounded sample of the headline workload: one probe per core x enough
    gallery prints to fill about ``budget_s`` seconds x all 13 variants, at the full map shape."""
    from oracle import reference_arm
    from src.shoeprint_image_retrieval import synth

    w = WORKLOAD
    cores = procs or os.cpu_count() or 1
    nvar = 1 + len(w["rotations"])
    sample_q = cores
    # ~0.2 s per pair-variant and process at this shape with every core busy (SURVEY.md section 6 measured 0.17 s on 8 cores)
    sample_g = max(1, min(16, int(budget_s / (0.2 * nvar))))
    gallery = synth.make_gallery(7, sample_g, w["C"], w["h"], w["w"])
    probes, pairs = synth.make_probes(8, gallery, sample_q, min_frac=1.0)
    if reference_arm.available():
        _, seconds, used = reference_arm.time_workers(probes, gallery, pairs, w["rotations"], w["scales"], cores)
        kind = "reference"
        how = "unmodified reference _comparison_worker (similarity.py:287-375) from baseline/_ref, one forked process per host core"
    else:
        from oracle import compare as ocmp

        _, seconds, used = ocmp.pair_variants_per_second(probes, gallery, w["rotations"], w["scales"], cores)
        kind = "port"
        how = "oracle port (baseline/_ref absent), one forked process per host core"
    return {
        "value": sample_q * sample_g / seconds,
        "unit": "pairs/s",
        "cores": used,
        "kind": kind,
        "sample": f"{sample_q} probes x {sample_g} gallery x {nvar} variants at C={w['C']} {w['h']}x{w['w']} ({sample_q * sample_g * nvar} pair-variants, {seconds:.1f} s); {how}",
        "seconds": seconds,
        "pairs": sample_q * sample_g,
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge

    with contextlib.suppress(Exception):
        ge.build()  # vendors baseline/_ref where /root/reference exists; the arm itself needs no GPU
    total = max(1, args.warmup + args.steps)
    budget = max(3.0, min(18.0, 200.0 / total))  # the whole run stays within a few minutes
    times, res = [], None
    for i in range(total):
        res = cpu_baseline(budget)
        if i >= args.warmup:
            times.append(res["seconds"])
    mean_s = sum(times) / len(times)
    value = res["pairs"] / mean_s
    cb = {k: res[k] for k in ("unit", "cores", "kind", "sample")}
    cb["value"] = value
    line = {
        "impl": "reference", "metric": "probe x gallery pairs scored/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 numerator / f64 window energy (scipy.signal.convolve)",
        "data": "synthetic", "config": _config(max(1, args.gpus), args.precision),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- feature stage (side measurement)
def feature_stage_parallel(world: int, rank: int, dist, torch) -> dict:
    """Feature stage data-parallel over the GPUs (images are independent): every rank pushes its own 64-image
    batches of 800x300 prints through the backbone, device-resident; aggregate images/s from the slowest rank."""
    import numpy as np

    from src.shoeprint_image_retrieval import network

    cfg = {"model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
    model = network.Model(cfg, 6, random_init_seed=0)
    rng = np.random.default_rng(100 + rank)
    d_batch = torch.from_numpy(rng.integers(0, 256, size=(64, 800, 300), dtype=np.uint8)).cuda()
    model._forward_device(d_batch, apply_clahe=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 4
    for _ in range(reps):
        model._forward_device(d_batch, apply_clahe=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return {"n_gpus": world, "images_per_s": world * reps * 64 / (float(ms.item()) * 1e-3), "per_gpu_batch": 64,
            "note": "device-resident uint8 prints -> [C,h,w] maps, CLAHE included; one process per GPU, no collective"}


def feature_stage_numbers(args) -> dict:
    """images/s of the backbone (EfficientNetV2-M cut at block 6, seeded random init, synthetic 800x300
    uint8 prints): through ``Model.get_multiple_feature_maps`` (H2D, GPU CLAHE, kernels, D2H) and
    device-only (CLAHE included); next to it the reference's own GPU path on this box (the torch modules on
    "cuda" through cuDNN, network.py:116,189,228-235: batch 1 as the reference runs it, and batch 64) and its CPU path."""
    import numpy as np
    import torch

    from src.shoeprint_image_retrieval import network

    cfg = {"model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
    model = network.Model(cfg, 6, random_init_seed=0)
    rng = np.random.default_rng(0)
    n = 512  # eight 64-image chunks: the API overlaps staging / result copies with the next chunk's kernels
    imgs = [np.clip(np.kron(rng.integers(0, 256, size=(100, 38)), np.ones((8, 8))) + rng.normal(0, 10, (800, 304)), 0, 255).astype(np.uint8)[:, :300] for _ in range(n)]
    imgs = [np.ascontiguousarray(im) for im in imgs]
    os.environ["SIR_FEATURE_CACHE"] = "0"  # time the kernels, not the cache
    model.get_multiple_feature_maps(imgs[:64], progress=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    maps = model.get_multiple_feature_maps(imgs, progress=False)
    torch.cuda.synchronize()
    e2e = n / (time.perf_counter() - t0)
    os.environ.pop("SIR_FEATURE_CACHE", None)
    d_batch = torch.from_numpy(np.stack(imgs[:64])).cuda()  # device-resident uint8 prints: CLAHE + backbone + layout change timed
    model._forward_device(d_batch, apply_clahe=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        model._forward_device(d_batch, apply_clahe=True)
    e1.record()
    torch.cuda.synchronize()
    dev = 2 * 64 / (e0.elapsed_time(e1) * 1e-3)
    out = {"unit": "images/s", "model": "EfficientNetV2_M[:6]", "input": "800x300 uint8", "map": list(maps[0].shape),
           "e2e_images_per_s": e2e, "device_images_per_s": dev, "gflop_per_image": 34.72,
           "device_tflops_algorithmic": dev * 34.72e-3,
           # whole stage (CLAHE, depthwise, SE, layout change included) against the measured sustained bf16 peak; the convolutions
           # spend three fp16 MMAs per algorithmic MAC (float32-grade hi/lo products), so the tensor-pipe share is 3x this
           "roofline_frac_algorithmic": dev * 34.72e-3 / _peaks()["bf16"], "mma_per_algorithmic_mac": 3}
    # the reference's own device path on this GPU: the same torch modules through cuDNN (default torch flags: TF32 convolutions allowed)
    try:
        net = model.model.to("cuda").eval()
        mean = torch.tensor(model.mean, device="cuda").view(1, 3, 1, 1)
        std = torch.tensor(model.std, device="cuda").view(1, 3, 1, 1)

        def torch_ips(batch: int, reps: int) -> float:
            x = ((d_batch[:batch].float() / 255.0).unsqueeze(1).repeat(1, 3, 1, 1) - mean) / std
            with torch.no_grad():
                for _ in range(3):
                    net(x)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    net(x)
                b.record()
                torch.cuda.synchronize()
            return batch * reps / (a.elapsed_time(b) * 1e-3)

        out["torch_cudnn_images_per_s"] = {"batch1": torch_ips(1, 20), "batch64": torch_ips(64, 3),
                                           "note": "torch modules on this B200 (network.py:228-235 runs batch 1), no CLAHE / copies, "
                                                   f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}"}
        model.model.to("cpu")
    except Exception as exc:  # a comparator must not take the bench down
        out["torch_cudnn_images_per_s"] = {"error": str(exc)[:200]}
    if not args.no_cpu:
        from oracle import features as ofeat

        ips, threads = ofeat.images_per_second(model.model, [model._clahe(im) for im in imgs[:6]], model.mean, model.std)
        out["cpu_images_per_s"] = ips
        out["cpu_threads"] = threads
    return out


# ----------------------------------------------------------------------------- GPU arm helpers
def _phase_ms(events: dict) -> dict:
    return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in events.items()}


def config4_block(args, world: int, rank: int, dist, torch) -> dict | None:
    """BASELINE configs[3]: 1,000 probes x (12,500 x N) on-device gallery maps of 80x59x21 -- the 1,000 x 100,000 target at
    N = 8 -- gallery sharded by contiguous index range, rank + top-64 merged over NCCL.  V = 1 and V = 13."""
    from src.shoeprint_image_retrieval import engine, sharding, synth

    c4 = CONFIG4
    q, g_local = c4["Q"], c4["G_per_gpu"]
    g_total, g0 = g_local * world, rank * g_local
    gal = synth.device_gallery(3000 + rank, g_local, c4["C"], c4["h"], c4["w"])
    q0, q1 = sharding.shard_range(q, world, rank)
    prb_local, pairs_local = synth.device_probes(4000 + rank, gal, q1 - q0)
    pairs_local = pairs_local + g0
    if world > 1:
        parts, pparts = [], []
        for r in range(world):
            a, b = sharding.shard_range(q, world, r)
            buf = prb_local if r == rank else torch.empty((b - a, *prb_local.shape[1:]), dtype=torch.float32, device="cuda")
            pb = pairs_local if r == rank else torch.empty(b - a, dtype=torch.int32, device="cuda")
            dist.broadcast(buf, src=r)
            dist.broadcast(pb, src=r)
            parts.append(buf)
            pparts.append(pb)
        prb, pairs = torch.cat(parts), torch.cat(pparts)
    else:
        prb, pairs = prb_local, pairs_local
    ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
    peaks = _peaks()
    out = {"workload": c4["name"], "Q": q, "G": g_total, "G_per_gpu": g_local, "top_k": c4["top_k"], "precision": args.precision,
           "note": "12,500 gallery maps per GPU at every N, so N = 8 is the 1,000 x 100,000 configuration of BASELINE configs[3]; "
                   "maps generated on the device (39.6 GB of float32 maps would not be a host list)"}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for label, rots, warm in (("v1", None, 1), ("v13", c4["rotations13"], 0)):
        if label == "v13" and args.quick:
            continue
        for _ in range(warm):
            sharding.compare_sharded(ps, gs, pairs, g0, rots, None, args.precision, c4["top_k"])
        engine.kernel_events, engine.refine_events, sharding.phase_events = [], [], {}
        launches0 = engine.launch_counter.n
        torch.cuda.reset_peak_memory_stats()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ranks, tv, ti, scores = sharding.compare_sharded(ps, gs, pairs, g0, rots, None, args.precision, c4["top_k"])
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        kev, rev, phases = engine.kernel_events, engine.refine_events, _phase_ms(sharding.phase_events)
        engine.kernel_events = engine.refine_events = sharding.phase_events = None
        k_ms = sum(a.elapsed_time(b) for a, b, _ in kev)
        k_fl = sum(f for _, _, f in kev)
        r_ms = sum(a.elapsed_time(b) for a, b in rev)
        nvar = 1 + (len(rots) if rots else 0)
        # correctness on hardware: the distributed ranks / top-k of 16 probes against a plain torch evaluation of their
        # gathered score rows (checks the count + candidate merge end to end)
        sub = 16
        rows = scores[:sub].contiguous()
        if world > 1:
            gathered = [torch.empty_like(rows) for _ in range(world)]
            dist.all_gather(gathered, rows)
            rows = torch.cat(gathered, dim=1)
        ts = rows.gather(1, pairs[:sub].long().unsqueeze(1))
        want = 1 + (rows > ts).sum(1)
        ok_ranks = bool(torch.equal(want.to(torch.int32), ranks[:sub].to(torch.int32)))
        ok_topk = bool(torch.equal(torch.topk(rows, c4["top_k"], dim=1).values, tv[:sub]))
        out[label] = {
            "variants": nvar, "pairs_per_s": q * g_total / (ms * 1e-3), "ms": ms,
            "screen_kernel": {"ms": k_ms, "tflops_algorithmic": k_fl / (k_ms * 1e-3) / 1e12 if k_ms else None,
                              "frac_of_sustained_bf16": (k_fl / (k_ms * 1e-3) / 1e12 / peaks["bf16"]) if k_ms else None, "share_of_pass": k_ms / ms},
            "refine_kernel_ms": r_ms,
            "prepare_ms": phases.get("scores", 0.0) - k_ms - r_ms,  # gallery / template pack, variants, window norms
            "merge_ms": {k: phases.get(k, 0.0) for k in ("true_score", "allreduce_max", "rank_topk", "allreduce_sum", "allgather", "merge_topk")},
            "rank_skew_ms": phases.get("rank_skew", 0.0),
            "rank1_share": float((ranks == 1).float().mean().item()),
            "ranks_match_gathered_rows": ok_ranks, "topk_matches_gathered_rows": ok_topk,
            "peak_device_gib": torch.cuda.max_memory_allocated() / 2**30, "gpu_launches": engine.launch_counter.n - launches0,
        }
        assert ok_ranks and ok_topk, f"config4 {label}: distributed ranks / top-k differ from the gathered rows"
    return out if rank == 0 else None


def config0_block(args, which: int = 0) -> dict:
    """BASELINE configs[0] / [2] shape of work through ``similarity.compare_maps`` from host lists -- the multi-shape bucket
    path (sir_template_pack_screen with a bucket layout + one window-norm table per 8 columns + per-tile row ranges).

    which = 0: the default run.toml on a FID-300-shaped set -- 300 RAGGED probes (every probe its own template shape, crops of
    40-100 % of a print) x 1,175 gallery maps of 80x59x21 (block 4) x 25 variants (7 rotations and 3 scales in the
    reference's combination).
    which = 2: the deeper cut of configs[2] -- block 6 on the same images, 176x30x11 maps, 12 rotations (+-30 step 5) and 3
    scales = 40 variants."""
    import torch

    from src.shoeprint_image_retrieval import engine, similarity, synth

    if which == 0:
        q, g, c, h, w = 300, 1175, 80, 59, 21
        rot, scl = [-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08]
        name = "configs[0]: default run.toml, FID-300 shape, 300 ragged probes x 1,175 gallery x 25 variants, host lists through compare_maps"
    else:
        q, g, c, h, w = 300, 1175, 176, 30, 11
        rot, scl = [a for a in range(-30, 31, 5) if a], [1.02, 1.04, 1.08]
        name = ("configs[2]: FID-300 shape at block 6 (176x30x11 maps), 300 ragged probes x 1,175 gallery x 40 variants "
                "(12 rotations x 3 scales + the unscaled set), host lists through compare_maps")
    gallery = synth.make_gallery(1, g, c, h, w)
    probes, pairs = synth.make_probes(2, gallery, q, min_frac=0.4 if which == 0 else 0.6)
    cfg = {"comparison": {"n_processes": 1, "rotations": rot, "scales": scl, "precision": args.precision}}
    # algorithmic work: 2*C*(gallery positions)*(template taps) per (variant, pair), template taps of the variant's true shape
    taps = 0
    plan = engine.variant_plan(rot, scl)
    for pm in probes:
        for r, sc in plan:
            hh, ww = pm.shape[1:]
            if sc is not None:
                hh, ww = engine.scaled_size(hh, ww, sc)
            taps += (hh - 4) * (ww - 4)
    flops = 2.0 * c * (h - 4) * (w - 4) * taps * g
    times = []
    for _ in range(2):
        sink = io.StringIO()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            ranks = similarity.compare_maps(probes, gallery, pairs, cfg)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    return {"workload": name, "variants": len(plan),
            "pairs_per_s": q * g / dt, "seconds": dt, "tflops_algorithmic_whole_pass": flops / dt / 1e12,
            "template_shapes": len({p_.shape for p_ in probes}), "rank1_share": float((ranks == 1).mean())}


def precision_block(args, world: int, rank: int, dist, torch) -> dict | None:
    """BASELINE configs[4]: 176x68x132 maps (1024x2048 inputs), 64 probes x 1,250 gallery maps per GPU (10,000 over 8).
    Every precision mode: pairs/s and max / mean relative score error against the float32 CUDA-core evaluation on a
    64 x 64 pair subsample, rank agreement, plus the CPU oracle (float64 restatement of similarity.py:26-108) on 4 x 4 pairs."""
    import numpy as np

    from src.shoeprint_image_retrieval import engine, synth

    c5 = CONFIG5
    q, g_local = c5["Q"], c5["G_per_gpu"] if not args.quick else 64
    gal = synth.device_gallery(5000 + rank, g_local, c5["C"], c5["h"], c5["w"])
    prb, pairs = synth.device_probes(6000 + rank, gal[:64], q, noise=0.6)
    ps, gs = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
    sub_g = engine.MapSet.from_device(gal[:64].contiguous())
    exact = engine.score_matrix(ps, sub_g, None, None, "fp32_simt")
    out = {"workload": c5["name"], "Q": q, "G": g_local * world, "G_per_gpu": g_local, "variants": 1,
           "reference": "float32 CUDA-core evaluation of the whole surface (fp32_simt) on the first 64 gallery maps; cpu_oracle on 4 x 4 pairs",
           "note": "bf16 / tf32 operands are not built: fp16 has the same 11-bit significand as tf32 at twice its tensor rate and 3 more bits "
                   "than bf16, so fp16x1 is the lossy single-pass point of the study and fp16_refine / fp16_fp8c / fp16x3 the compensated ones",
           "modes": {}}
    for mode in ("fp16_refine", "fp16_fp8c", "fp16x3", "fp16x1"):
        got = engine.score_matrix(ps, sub_g, None, None, mode)
        rel = ((got - exact).abs() / exact.abs().clamp_min(1e-3))
        agree = float((got.argmax(1) == exact.argmax(1)).float().mean().item())
        entry = {"max_rel_err": float(rel.max().item()), "mean_rel_err": float(rel.mean().item()), "top1_agreement_with_fp32": agree}
        if mode == args.precision or not args.quick:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.score_matrix(ps, gs, None, None, mode)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            entry["pairs_per_s"] = q * g_local * world / (float(ms.item()) * 1e-3)
            entry["tflops_algorithmic"] = 2.0 * c5["C"] * ((c5["h"] - 4) * (c5["w"] - 4)) ** 2 * entry["pairs_per_s"] / 1e12
        out["modes"][mode] = entry
    if rank == 0 and not args.no_cpu:
        from oracle import compare as ocmp

        hp = [prb[i].cpu().numpy() for i in range(4)]
        hg = [gal[i].cpu().numpy() for i in range(4)]
        _, want = ocmp.compare_maps_oracle(hp, hg, [0] * 4, None, None)
        got = engine.score_matrix(engine.MapSet.from_device(prb[:4].contiguous()), engine.MapSet.from_device(gal[:4].contiguous()), None, None, args.precision)
        err = np.abs(got.cpu().numpy() - want) / np.maximum(np.abs(want), 1e-3)
        out["cpu_oracle_4x4"] = {"mode": args.precision, "max_rel_err": float(err.max()), "tolerance": 1e-4}
    return out if rank == 0 else None


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from src.shoeprint_image_retrieval import engine, sharding, similarity, synth

    w = dict(WORKLOAD)
    if args.profile:  # shorter kernel for ncu replays; same shapes
        w["Q"] = 256
    q_total, g_local = w["Q"], w["G_per_gpu"]
    g_total = g_local * world
    g0 = rank * g_local
    nvar = 1 + len(w["rotations"])

    # data: every rank owns a gallery shard and makes the probes whose match lives on it
    gal = synth.device_gallery(1000 + rank, g_local, w["C"], w["h"], w["w"])
    q0, q1 = sharding.shard_range(q_total, world, rank)
    prb_local, pairs_local = synth.device_probes(2000 + rank, gal, q1 - q0)
    pairs_local = pairs_local + g0
    if world > 1:
        sizes = [sharding.shard_range(q_total, world, r) for r in range(world)]
        parts, pparts = [], []
        for r, (a, b) in enumerate(sizes):
            buf = prb_local if r == rank else torch.empty((b - a, *prb_local.shape[1:]), dtype=torch.float32, device="cuda")
            pb = pairs_local if r == rank else torch.empty(b - a, dtype=torch.int32, device="cuda")
            dist.broadcast(buf, src=r)
            dist.broadcast(pb, src=r)
            parts.append(buf)
            pparts.append(pb)
        prb, pairs = torch.cat(parts), torch.cat(pparts)
    else:
        prb, pairs = prb_local, pairs_local
    probes_dev, gallery_dev = engine.MapSet.from_device(prb), engine.MapSet.from_device(gal)
    # host inputs of the e2e leg: what a caller of the reference API holds -- one ordinary (pageable) float32 numpy array
    # per feature map, as Model.get_multiple_feature_maps returns them.  Under torchrun every rank passes the same lists
    # (SPMD); the entries of other ranks' gallery shards are never read, so they all alias one dummy array.
    prb_host = prb.cpu().numpy()
    probes_host = [np.array(prb_host[i]) for i in range(prb_host.shape[0])]
    gal_host = gal.cpu().numpy()
    dummy = np.zeros(gal_host.shape[1:], dtype=np.float32)
    gallery_host = [dummy] * g0 + [np.array(gal_host[i]) for i in range(g_local)] + [dummy] * (g_total - g0 - g_local)
    pairs_host = [int(v) for v in pairs.cpu().tolist()]
    api_config = {"comparison": {"n_processes": 1, "rotations": w["rotations"], "scales": w["scales"], "precision": args.precision, "top_k": w["top_k"]}}

    def step_device():
        return sharding.compare_sharded(probes_dev, gallery_dev, pairs, g0, w["rotations"], w["scales"], args.precision, w["top_k"])

    def step_e2e():
        sink = io.StringIO()  # compare_maps prints one "Print i true match ranked r" line per probe (similarity.py:375)
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            ranks = similarity.compare_maps(probes_host, gallery_host, pairs_host, api_config)
        return ranks, sum(similarity.last_result["h2d_bytes"])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(max(n_warm, 1)):
        out = step_device()
    ranks_dev = out[0]

    # device-resident timing (+ per-kernel events for the roofline)
    sampler = ClockSampler(local)
    engine.kernel_events, engine.refine_events = [], []
    engine.collect_refine_stats = True
    engine.refine_stats(reset=True)
    launches0 = engine.launch_counter.n
    sampler.start()
    ms_total, _ = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches = engine.launch_counter.n - launches0
    kev, rev = engine.kernel_events, engine.refine_events
    rstats = engine.refine_stats(reset=True)
    engine.kernel_events = engine.refine_events = None
    engine.collect_refine_stats = False
    k_ms = [a.elapsed_time(b) for a, b, _ in kev]
    k_flops = [f for _, _, f in kev]
    r_ms = [a.elapsed_time(b) for a, b in rev]
    ms_step = ms_total / args.steps

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_step, "kernel_ms": k_ms, "refine_ms": r_ms}), flush=True)
        return
    # end-to-end timing through the reference-facing API from pageable host lists, at least 10 steps
    step_e2e()
    e2e_steps = max(10, args.steps) if not args.quick else 2
    t_e2e, out_e2e = timed(step_e2e, e2e_steps)
    ms_e2e = t_e2e / e2e_steps
    ranks_host, h2d = out_e2e
    acc = float((ranks_dev == 1).float().mean().item())
    assert np.array_equal(np.asarray(ranks_host, dtype=np.int32), ranks_dev.cpu().numpy().astype(np.int32)), "e2e and device-resident ranks differ"
    del probes_host, gallery_host, prb_host, gal_host

    c4 = config4_block(args, world, rank, dist, torch) if not args.no_config4 else None
    c5 = precision_block(args, world, rank, dist, torch) if not args.no_precision_study else None
    c0 = config0_block(args) if (rank == 0 and world == 1 and not args.no_config4 and not args.quick) else None
    c2 = config0_block(args, 2) if c0 is not None else None
    feat_par = feature_stage_parallel(world, rank, dist, torch) if (world > 1 and not args.no_features) else None
    feat = feature_stage_numbers(args) if (rank == 0 and not args.no_features) else None
    if feat is not None and feat_par is not None:
        feat["data_parallel"] = feat_par

    if rank == 0:
        peaks = _peaks()
        achieved = sum(k_flops) / (sum(k_ms) * 1e-3) / 1e12 if k_ms else None
        pairs_per_step = q_total * g_total
        cb = cpu_baseline() if world == 1 and not args.no_cpu else None
        if cb:
            cb.pop("pairs", None)
        traffic, traffic_src = _ncu_traffic()
        line = {
            "metric": "probe x gallery pairs scored/sec",
            "value": pairs_per_step / (ms_step * 1e-3),
            "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPES[args.precision],
            "data": "synthetic",
            "config": _config(world, args.precision),
            "e2e": {"value": pairs_per_step / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(len(ranks_host) * 4),
                    "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "api": "similarity.compare_maps(list[np.ndarray], list[np.ndarray], list[int], config) -> np.ndarray[int32]; pageable host arrays"},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "ncc_tc_kernel", "bound": "tensor", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["bf16"]) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["src"],
                "launches": len(k_ms), "mean_launch_ms": (sum(k_ms) / len(k_ms)) if k_ms else None,
                "share_of_step": (sum(k_ms) / ms_total) if k_ms else None,
                "note": "achieved = algorithmic 2*C*M*K FLOPs per (column, gallery) / event time of the screening launches: one fp16 MMA per "
                        "algorithmic MAC; tile padding (+19%) is not counted, the K steps that multiply only zero padding (12%) are skipped",
            },
            "refine": {"kernel": "ncc_refine_kernel", "ms_per_step": sum(r_ms) / args.steps if r_ms else None,
                       "share_of_step": (sum(r_ms) / ms_total) if r_ms else None,
                       "positions_per_pair": rstats["positions"] / (args.steps * q_total * g_local) if rstats["positions"] else None,
                       "dense_records_per_step": rstats["dense_records"] / args.steps},
            "cpu_baseline": cb,
            "config0": c0,
            "config2": c2,
            "config4": c4,
            "precision_study": c5,
            "feature_stage": feat,
            "clocks": clocks,
            "top1_accuracy": acc,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16_refine", choices=PRECISIONS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the CPU comparators")
    ap.add_argument("--no-features", action="store_true", help="skip the feature-stage side measurement")
    ap.add_argument("--no-config4", action="store_true", help="skip the configs[3] gallery-scaling block")
    ap.add_argument("--no-precision-study", action="store_true", help="skip the configs[4] precision-study block")
    ap.add_argument("--quick", action="store_true", help="development: shorter side blocks (no V=13 pass, fewer e2e steps)")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 256 probes, no e2e / cpu legs, warm-up as given")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
