#!/usr/bin/env python3
"""Run under torchrun on >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Gallery-sharded compare (NCCL merge) must equal the single-GPU result on identical inputs: scores are
computed by the same kernels per shard, so ranks and the merged top-k must match exactly (SURVEY.md
section 4 iv: sharding must not change any arithmetic)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as ge

if rank == 0:
    ge.build()
dist.barrier()
from src.shoeprint_image_retrieval import engine, sharding, synth

g_total, q, k = 61, 23, 7  # odd sizes on purpose
gal = synth.device_gallery(5, g_total, 12, 30, 21)  # same seed on every rank -> identical full gallery
prb, pairs = synth.device_probes(6, gal, q)
rot, scl = [-7, 7], [1.05]
g0, g1 = sharding.shard_range(g_total, world, rank)
ranks, tv, ti, _ = sharding.compare_sharded(engine.MapSet.from_device(prb), engine.MapSet.from_device(gal[g0:g1].contiguous()),
                                            pairs, g0, rot, scl, k=k)
if rank == 0:
    full = engine.score_matrix(engine.MapSet.from_device(prb), engine.MapSet.from_device(gal), rot, scl)
    gt, _, fv, fi, _ = engine.rank_true_matches(full, pairs, k)
    assert torch.equal(ranks, (gt + 1).to(torch.int32)), (ranks, gt + 1)
    assert torch.equal(ti, fi), (ti, fi)
    assert torch.equal(tv, fv)
    print(f"multi-GPU check ok on {world} GPUs: ranks and top-{k} identical to the single-GPU pass; ranks {ranks.tolist()}")
dist.barrier()
dist.destroy_process_group()
