#!/usr/bin/env python3
"""Benchmark of the matching hot path: probe x gallery pairs scored per second.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp16x3]

Workload (BASELINE.json configs[1]): synthetic WVU2019-shaped set -- Q = 1,500 probes x
G = 150 gallery prints PER GPU, feature maps of the default backbone cut at block 6 on 800x300
images (C = 176, 50 x 19), rotation sweep -30..30 step 5 degrees (12 angles + the unrotated probe =
13 variants).  A *step* is one full compare pass over that batch: gallery pack, variant
generation, template pack, window norms, correlation over all variants with the fused max, rank
of the true match and top-k (+ the NCCL merge when N > 1).  A pair = one (probe, gallery print),
all its variants included.

* ``value``  pairs/s with the float32 feature maps already resident in HBM (CUDA events, max over ranks)
* ``e2e``    the same pass through the public API ``engine.compare`` / ``compare_sharded`` from
             pinned-host feature-map lists to ranks on the host, H2D and D2H inside the timed region
* ``roofline`` for the dominant kernel (``ncc_tc_kernel``): algorithmic FLOPs / mean launch time
* ``cpu_baseline`` the oracle port (the reference's algorithm: three FFT convolutions per channel
             per pair) on the box's host cores, on a bounded sample of the same workload

Multi-GPU: one process per GPU (torchrun), gallery sharded by contiguous index range, 150 prints
per GPU (weak scaling); probes replicated; two tiny all-reduces + one all-gather per step.
"""

from __future__ import annotations

import argparse
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

# ----------------------------------------------------------------------------- workload
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
CPU_SAMPLE = {"Q": 16, "G": 16}  # bounded CPU sample: 16 x 16 pairs x 13 variants at the full map shape (~15-20 s on 8 cores)


def _ncu_traffic() -> tuple[float | None, str | None]:
    """DRAM bytes of one ncc_tc_kernel launch from the committed `ncu --set full` capture (profiles/)."""
    f = ROOT / "profiles" / "r01_ncu_full_ncc_tc_kernel_final.txt"
    if not f.exists():
        return None, None
    total = 0.0
    for line in f.read_text().splitlines():
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if line.startswith(key + " ["):
                unit = line.split("[")[1].split("]")[0]
                val = float(line.split("=")[1])
                total += val * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
    return total, "ncu --set full capture of the bench --profile launch (3,328 columns x 150 gallery; the bench's own launches are 16,384 and 3,116 columns wide and move proportionally more)"


def _peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"bf16": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "src": "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a ~1 s step)"}
    return {"bf16": 1400.0, "src": "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"}


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
def cpu_baseline(sample_q: int, sample_g: int, procs: int | None = None) -> dict:
    """The oracle port timed on the host cores over a bounded sample of the workload."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    w = WORKLOAD
    gallery = synth.make_gallery(7, sample_g, w["C"], w["h"], w["w"])
    probes, _ = synth.make_probes(8, gallery, sample_q, min_frac=1.0)
    _, seconds, used = ocmp.pair_variants_per_second(probes, gallery, w["rotations"], w["scales"], procs)
    nvar = 1 + len(w["rotations"])
    return {
        "value": sample_q * sample_g / seconds,
        "unit": "pairs/s",
        "cores": used,
        "kind": "port",
        "sample": f"{sample_q} probes x {sample_g} gallery x {nvar} variants at C={w['C']} {w['h']}x{w['w']} ({sample_q * sample_g * nvar} pair-variants, {seconds:.1f} s)",
        "seconds": seconds,
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(CPU_SAMPLE["Q"], CPU_SAMPLE["G"])
        if i >= args.warmup:
            times.append(res["seconds"])
    mean_s = sum(times) / len(times)
    value = CPU_SAMPLE["Q"] * CPU_SAMPLE["G"] / mean_s
    cb = {k: res[k] for k in ("unit", "cores", "kind", "sample")}
    cb["value"] = value
    line = {
        "impl": "reference", "metric": "probe x gallery pairs scored/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (scipy-style FFT)",
        "data": "synthetic", "config": {"workload": WORKLOAD["name"], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- feature stage (side measurement)
def feature_stage_numbers(args) -> dict:
    """images/s of the backbone (EfficientNetV2-M cut at block 6, seeded random init, synthetic 800x300
    uint8 prints): through ``Model.get_multiple_feature_maps`` (H2D, GPU CLAHE, kernels, D2H) and
    device-only (CLAHE included); plus the torch CPU forward of the same modules (the reference's path on a CPU box)."""
    import numpy as np
    import torch

    from src.shoeprint_image_retrieval import network

    cfg = {"model": {"type": "EfficientNetV2_M", "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}
    model = network.Model(cfg, 6, random_init_seed=0)
    rng = np.random.default_rng(0)
    n = 512  # eight 64-image chunks: the API overlaps staging / result copies with the next chunk's kernels
    imgs = [np.clip(np.kron(rng.integers(0, 256, size=(100, 38)), np.ones((8, 8))) + rng.normal(0, 10, (800, 304)), 0, 255).astype(np.uint8)[:, :300] for _ in range(n)]
    imgs = [np.ascontiguousarray(im) for im in imgs]
    model.get_multiple_feature_maps(imgs[:64], progress=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    maps = model.get_multiple_feature_maps(imgs, progress=False)
    torch.cuda.synchronize()
    e2e = n / (time.perf_counter() - t0)
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
    if not args.no_cpu:
        from oracle import features as ofeat

        ips, threads = ofeat.images_per_second(model.model, [model._clahe(im) for im in imgs[:6]], model.mean, model.std)
        out["cpu_images_per_s"] = ips
        out["cpu_threads"] = threads
    return out


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
    from src.shoeprint_image_retrieval import engine, sharding, synth

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
    # host inputs of the e2e leg: one numpy array per map (the reference API's lists), living in pinned memory
    prb_pin, gal_pin = prb.cpu().pin_memory(), gal.cpu().pin_memory()
    probes_host = [prb_pin[i].numpy() for i in range(prb_pin.shape[0])]
    gallery_host = [gal_pin[i].numpy() for i in range(gal_pin.shape[0])]
    pairs_host = pairs.cpu().tolist()

    def step_device():
        return sharding.compare_sharded(probes_dev, gallery_dev, pairs, g0, w["rotations"], w["scales"], args.precision, w["top_k"])

    def step_e2e():
        ps = engine.MapSet.from_host(probes_host)
        gs = engine.MapSet.from_host(gallery_host)
        ranks, _, _, _ = sharding.compare_sharded(ps, gs, pairs_host, g0, w["rotations"], w["scales"], args.precision, w["top_k"])
        return ranks.cpu(), ps.h2d_bytes + gs.h2d_bytes

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
    engine.kernel_events = []
    launches0 = engine.launch_counter.n
    sampler.start()
    ms_total, _ = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches = engine.launch_counter.n - launches0
    kev = engine.kernel_events
    engine.kernel_events = None
    k_ms = [a.elapsed_time(b) for a, b, _ in kev]
    k_flops = [f for _, _, f in kev]
    ms_step = ms_total / args.steps

    # end-to-end timing through the public API from host buffers
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_step, "kernel_ms": k_ms}), flush=True)
        return
    step_e2e()
    t_e2e, out_e2e = timed(step_e2e, max(1, min(args.steps, 3)))
    ms_e2e = t_e2e / max(1, min(args.steps, 3))
    ranks_host, h2d = out_e2e

    feat = feature_stage_numbers(args) if (rank == 0 and not args.no_features) else None
    acc = float((ranks_dev == 1).float().mean().item())
    assert torch.equal(ranks_host.to(torch.int32), ranks_dev.cpu().to(torch.int32)), "e2e and device-resident ranks differ"

    if rank == 0:
        peaks = _peaks()
        achieved = sum(k_flops) / (sum(k_ms) * 1e-3) / 1e12 if k_ms else None
        pairs_per_step = q_total * g_total
        cb = cpu_baseline(CPU_SAMPLE["Q"], CPU_SAMPLE["G"]) if world == 1 and not args.no_cpu else None
        line = {
            "metric": "probe x gallery pairs scored/sec",
            "value": pairs_per_step / (ms_step * 1e-3),
            "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16x3": "fp16 hi/lo split x3 MMAs, fp32 accumulate (fp32-grade)", "fp16x1": "fp16, fp32 accumulate", "fp32_simt": "fp32", "fp16_fp8c": "fp16 hi*hi + fp8 e4m3 correction MMAs, fp32 accumulate (fp32-grade within 1e-4)"}[args.precision],
            "data": "synthetic",
            "config": {
                "workload": w["name"], "Q": q_total, "G": g_total, "G_per_gpu": g_local, "C": w["C"],
                "map_hw": [w["h"], w["w"]], "variants": nvar, "top_k": w["top_k"], "precision": args.precision,
                "sharding": f"gallery x{world}" if world > 1 else "none",
                "l2": "inputs (1.0 GB of probe maps + 13 variants) exceed the 126 MB L2; no explicit flush",
            },
            "e2e": {"value": pairs_per_step / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(ranks_host.numel() * 4), "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "ncc_tc_kernel", "bound": "tensor", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["bf16"]) if achieved else None,
                "traffic": _ncu_traffic()[0], "traffic_source": _ncu_traffic()[1],
                "peak_source": peaks["src"],
                "launches": len(k_ms), "mean_launch_ms": (sum(k_ms) / len(k_ms)) if k_ms else None,
                "share_of_step": (sum(k_ms) / ms_total) if k_ms else None,
                "note": "achieved = algorithmic 2*C*M*K FLOPs per (column, gallery) / event time. Per algorithmic MAC the kernel spends 2 fp16-MMA-equivalents in fp16_fp8c (3 in fp16x3) and tile padding adds ~19%; neither is counted, 12% of the K steps multiply only zero padding and are skipped",
            },
            "cpu_baseline": cb,
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
    ap.add_argument("--precision", default="fp16_fp8c", choices=["fp16x3", "fp16x1", "fp32_simt", "fp16_fp8c"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-features", action="store_true", help="skip the feature-stage side measurement")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 256 probes, no e2e / cpu legs, warm-up as given")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
