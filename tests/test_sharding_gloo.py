"""N>1 host logic on CPU: world_size-2 gloo run of the sharded rank merge (sharding.py).

The per-shard kernels (scores, sir_rank_topk, sir_merge_topk) are CUDA and are covered by the -m gpu
tests; here the local halves are emulated with numpy so that the COLLECTIVE plumbing (shard ranges,
max all-reduce of the true score, count reduction, candidate all-gather + merge order) is what is
under test.  The torch merge used as ``merge_fn`` lives in this test file only."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torch_merge(all_v, all_i, k):
    p, q, _ = all_v.shape
    v = all_v.permute(1, 0, 2).reshape(q, p * k)
    i = all_i.permute(1, 0, 2).reshape(q, p * k).to(torch.int64)
    key_i = torch.where(i < 0, torch.full_like(i, 2**31 - 1), i)
    order = np.lexsort((key_i.numpy(), -v.numpy()), axis=1)[:, :k]
    order = torch.from_numpy(order)
    return torch.gather(v, 1, order), torch.gather(i, 1, order).to(torch.int32)


def _worker(rank, world, port, scores, true_idx, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from src.shoeprint_image_retrieval import sharding

    q, g = scores.shape
    g0, g1 = sharding.shard_range(g, world, rank)
    local = scores[:, g0:g1]
    ts = torch.full((q,), float("-inf"))
    for i in range(q):
        if g0 <= true_idx[i] < g1:
            ts[i] = float(local[i, true_idx[i] - g0])
    dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    lt = torch.from_numpy(local)
    gt = (lt > ts[:, None]).sum(1).to(torch.int32)
    ge = (lt >= ts[:, None]).sum(1).to(torch.int32)
    kk = min(k, g1 - g0)
    order = np.lexsort((np.arange(g1 - g0)[None].repeat(q, 0), -local), axis=1)[:, :kk]
    tv = torch.full((q, k), float("-inf"))
    ti = torch.full((q, k), -1, dtype=torch.int32)
    tv[:, :kk] = torch.from_numpy(np.take_along_axis(local, order, 1))
    ti[:, :kk] = torch.from_numpy(order + g0).to(torch.int32)
    ranks, rank_hi, mv, mi = sharding.merge_ranks(ts, gt, ge, tv, ti, merge_fn=_torch_merge)
    if rank == 0:
        out["ranks"], out["hi"], out["v"], out["i"] = ranks.numpy(), rank_hi.numpy(), mv.numpy(), mi.numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_everything():
    from src.shoeprint_image_retrieval import sharding

    for total in (0, 1, 7, 150, 1175, 100000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("g,k", [(37, 5), (8, 6)])
def test_two_rank_merge_equals_single_process(g, k):
    rng = np.random.default_rng(3)
    q = 9
    scores = rng.random((q, g)).astype(np.float32)
    scores[:, ::5] = scores[:, 1:2]  # exact ties across shards
    true_idx = [int(x) for x in rng.integers(0, g, size=q)]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), scores, true_idx, k, out), nprocs=2, join=True)
        ts = scores[np.arange(q), true_idx]
        np.testing.assert_array_equal(out["ranks"], 1 + (scores > ts[:, None]).sum(1))
        np.testing.assert_array_equal(out["hi"], (scores >= ts[:, None]).sum(1))
        order = np.lexsort((np.arange(g)[None].repeat(q, 0), -scores), axis=1)[:, :k]
        np.testing.assert_array_equal(out["i"], order)
        np.testing.assert_array_equal(out["v"], np.take_along_axis(scores, order, 1))
