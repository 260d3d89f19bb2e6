"""Gallery sharding across GPUs: one process per GPU, contiguous gallery index ranges, NCCL for
the (small) rank / top-k merge.

Every (probe, gallery) score is independent (similarity.py:357-367), so each rank scores ALL probes
against ITS gallery shard with the single-GPU path -- sharding does not change any arithmetic.
The exact rank of the true match needs its score on every shard, hence two tiny collectives:

1. ``all_reduce(MAX)`` of ``true_score[Q]`` (the owning shard contributes the value, others -inf);
2. ``all_reduce(SUM)`` of ``count_gt[Q]`` / ``count_ge[Q]`` and one ``all_gather`` of the per-shard
   ``[Q, k]`` (score, global index) lists, merged by ``sir_merge_topk``.

``rank = 1 + sum over shards of #{s > s_true}`` is ``_get_rank`` (similarity.py:378-386) up to the
order of exact ties.  The reference shards the other way (probes over processes, gallery in shared
memory, similarity.py:146-176); results are identical either way.
"""

from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

__all__ = ["shard_range", "merge_ranks", "compare_sharded", "phase_events"]

#: when set to a dict, every phase of the sharded rank appends a (start, end) CUDA-event pair under its name
#: ("scores", "true_score", "rank_skew", "allreduce_max", "rank_topk", "allreduce_sum", "allgather", "merge_topk"); bench.py reads it
phase_events: dict | None = None


class _Phase:
    """``with _Phase("name"):`` brackets a phase with CUDA events on the current stream when ``phase_events`` is set."""

    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self):
        if phase_events is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if phase_events is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            phase_events.setdefault(self.name, []).append((self.e0, e1))
        return False


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous ``[start, end)`` of ``total`` items owned by ``rank`` (first ``total % world``
    ranks get one extra item -- the reference's probe chunking rule, similarity.py:146-157)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _merge_topk_cuda(vals: torch.Tensor, idx: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    from . import _native as nat
    from .engine import launch_counter

    p, q, _ = vals.shape
    out_v = torch.empty((q, k), dtype=torch.float32, device=vals.device)
    out_i = torch.empty((q, k), dtype=torch.int32, device=vals.device)
    nat.check(
        nat.lib.sir_merge_topk(C.c_void_p(vals.data_ptr()), C.c_void_p(idx.data_ptr()), p, q, k,
                               C.c_void_p(out_v.data_ptr()), C.c_void_p(out_i.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
        "sir_merge_topk",
    )
    launch_counter.add()
    return out_v, out_i


def merge_ranks(true_score, count_gt, count_ge, topk_val, topk_idx, group=None, merge_fn=None):
    """Collective half of the sharded rank: returns (ranks, rank_hi, topk_val, topk_idx), all global.

    ``true_score`` must already be the all-reduced value when counts were computed against it; this
    function reduces the counts and merges the candidate lists.  ``merge_fn`` defaults to the CUDA
    merge kernel (tests on the gloo backend pass a torch implementation)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    k = int(topk_val.shape[1])
    if world > 1:
        with _Phase("allreduce_sum"):
            counts = torch.stack([count_gt, count_ge]).to(torch.int32)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
            count_gt, count_ge = counts[0], counts[1]
        if k > 0:
            q = topk_val.shape[0]
            with _Phase("allgather"):
                all_v = torch.empty((world, q, k), dtype=topk_val.dtype, device=topk_val.device)
                all_i = torch.empty((world, q, k), dtype=topk_idx.dtype, device=topk_idx.device)
                dist.all_gather(list(all_v.unbind(0)), topk_val.contiguous(), group=group)
                dist.all_gather(list(all_i.unbind(0)), topk_idx.contiguous(), group=group)
            with _Phase("merge_topk"):
                topk_val, topk_idx = (merge_fn or _merge_topk_cuda)(all_v, all_i, k)
    return count_gt + 1, torch.clamp(count_ge, min=1), topk_val, topk_idx


def compare_sharded(probes, gallery_shard, true_idx, g0: int, rotations=None, scales=None,
                    precision: str = "fp16_refine", k: int = 0, group=None, packed_gallery=None):
    """Sharded compare pass on this rank: ``probes`` are replicated, ``gallery_shard`` holds global
    gallery indices ``[g0, g0 + G_local)``, ``true_idx`` are GLOBAL gallery indices.

    Returns (ranks int32 [Q] device, topk_val, topk_idx, local scores)."""
    from . import _native as nat
    from . import engine

    with _Phase("scores"):
        scores = engine.score_matrix(probes, gallery_shard, rotations, scales, precision, packed_gallery=packed_gallery)
    q, g = scores.shape
    dev = scores.device
    tidx = torch.as_tensor(true_idx, dtype=torch.int32).to(dev)
    with _Phase("true_score"):
        true_score = torch.empty(q, dtype=torch.float32, device=dev)
        nat.check(
            nat.lib.sir_true_scores(C.c_void_p(scores.data_ptr()), q, g, int(scores.stride(0)), C.c_void_p(tidx.data_ptr()), g0,
                                    C.c_void_p(true_score.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "sir_true_scores",
        )
        engine.launch_counter.add()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if phase_events is not None:
            # measuring only: the first collective of the merge otherwise absorbs the wait for the slowest rank's scoring
            # pass (tens of ms of clock skew between GPUs at 1,000 x 12,500 per rank); time that wait on its own
            with _Phase("rank_skew"):
                dist.all_reduce(torch.zeros(1, dtype=torch.float32, device=dev), group=group)
        with _Phase("allreduce_max"):
            dist.all_reduce(true_score, op=dist.ReduceOp.MAX, group=group)
    with _Phase("rank_topk"):
        gt, ge, tv, ti, _ = engine.rank_true_matches(scores, tidx, k, g0=g0, true_score=true_score)
    ranks, _, tv, ti = merge_ranks(true_score, gt, ge, tv, ti, group=group)
    return ranks.to(torch.int32), tv, ti, scores
