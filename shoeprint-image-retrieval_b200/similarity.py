"""Feature-map comparison on B200: drop-in for the reference ``similarity.py``.

Public names, arguments and results follow the reference (``similarity.py:26-386``):

* ``compare_maps(shoemark_maps, shoeprint_maps, matching_pairs, config) -> ndarray[int32]``
  1-based rank of each shoemark's true shoeprint (``similarity.py:129-227``);
* ``get_similarity(shoemark, shoeprint) -> float`` (``similarity.py:75-108``);
* ``normxcorr(template, image, mode="same") -> ndarray`` (``similarity.py:26-72``);
* ``MultiProcessingTrackers`` kept importable (``similarity.py:111-126``).

The work happens on the GPU through ``engine`` / ``libsir.so``; ``comparison.n_processes`` is
accepted and ignored (one process drives one GPU).  Under ``torchrun`` (``torch.distributed`` initialised with
more than one rank) the call is SPMD: every rank passes the same lists, scores ALL shoemarks against its contiguous
share of the shoeprints (``sharding.shard_range``), and the ranks are merged over NCCL (``sharding.compare_sharded``);
every rank returns the same result.  Differences a caller can observe:

* with both ``rotations`` and ``scales`` set the reference's progress loop never terminates
  (it waits for ``(R+1)(S+1)*Q`` ticks while the workers produce ``(1+(R+1)S)*Q``; SURVEY.md
  Appendix D2) -- this implementation scores the same variant set the workers score and returns;
* exact score ties are ranked best-first (``1 + #{s > s_true}``) instead of numpy's unspecified
  argsort order.
"""

from __future__ import annotations

from multiprocessing import Array, Queue, Value
from typing import TYPE_CHECKING, Literal

import numpy as np
from tqdm import tqdm

from . import engine

if TYPE_CHECKING:
    from .config import Config
    from .customtypes import FeatureMapsArrayType, ImageArrayType

_PAD = engine.EDGE

#: scores ``[Q, G]`` (torch, on the device), top-k lists and host->device bytes (``h2d_bytes``: (shoemarks,
#: shoeprints); 0 when the maps came from ``Model.get_multiple_feature_maps`` and were still resident on the
#: device) of the most recent ``compare_maps`` call
last_result: dict = {}


class MultiProcessingTrackers:
    """Shared progress state of the reference's worker processes (``similarity.py:111-126``).
    Unused by the GPU path; constructed on demand for callers that poke at it."""

    def __init__(self, rankings_len: int) -> None:
        self.counter = Value("i", 0)
        self.queue: Queue = Queue()
        self.rankings = Array("i", rankings_len)


def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _ingest(maps):
    """The caller's list, untouched when it can be used as is: a ``network.FeatureMapList`` whose device
    copies are still valid is handed through as the very same object (``engine`` then reads the maps where
    the feature stage left them, no upload); in any other list only the elements that are not already
    C-contiguous float32 arrays are converted."""
    copies = getattr(maps, "device_copies", None)
    if copies is not None and copies() is not None:
        return maps
    return [m if isinstance(m, np.ndarray) and m.dtype == np.float32 and m.flags.c_contiguous else _as_f32(m) for m in maps]


def compare_maps(
    shoemark_maps: list[FeatureMapsArrayType],
    shoeprint_maps: list[FeatureMapsArrayType],
    matching_pairs: list[int],
    config: Config,
) -> np.ndarray:
    """Rank every shoemark's true shoeprint among all shoeprints by NCC similarity.

    Args:
        shoemark_maps: probe feature maps, each ``[C, h, w]`` float32 (shapes may differ).
        shoeprint_maps: gallery feature maps, each ``[C, h, w]`` float32.
        matching_pairs: ``matching_pairs[i]`` = index into ``shoeprint_maps`` of shoemark ``i``'s match.
        config: system config; ``config["comparison"]`` supplies ``rotations`` / ``scales``.
    """
    comparison = config["comparison"]
    rotations = comparison.get("rotations")
    scales = comparison.get("scales")
    precision = comparison.get("precision", engine.DEFAULT_PRECISION)
    top_k = int(comparison.get("top_k", 0))
    if len(matching_pairs) < len(shoemark_maps):
        raise IndexError("matching_pairs is shorter than shoemark_maps")
    pairs = [int(v) for v in matching_pairs[: len(shoemark_maps)]]
    if any(v < 0 or v >= len(shoeprint_maps) for v in pairs):  # the reference fails in _get_rank (similarity.py:386)
        raise IndexError("matching_pairs holds an index outside shoeprint_maps")

    n_variants = len(engine.variant_plan(rotations, scales))
    world = rank = 0
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
    except ImportError:  # pragma: no cover
        pass
    with tqdm(total=n_variants * len(shoemark_maps), disable=rank != 0) as pbar:
        if world > 1:
            from . import sharding

            g0, g1 = sharding.shard_range(len(shoeprint_maps), world, rank)
            if g1 <= g0:
                raise ValueError(f"{len(shoeprint_maps)} shoeprints cannot be shared among {world} ranks")
            shard = engine.MapSet.from_host(_ingest(shoeprint_maps[g0:g1]), lazy=True)
            probes = engine.MapSet.from_host(_ingest(shoemark_maps), chunk=engine.PROBE_CHUNK, lazy=True)
            engine.last_h2d_bytes = (probes.h2d_bytes, shard.h2d_bytes)
            d_ranks, tv, ti, scores = sharding.compare_sharded(probes, shard, pairs, g0, rotations, scales, precision, top_k)
            ranks = d_ranks.to("cpu").numpy().astype(np.int32)
            topk = (tv, ti)
        else:
            ranks, scores, topk = engine.compare(
                _ingest(shoemark_maps),
                _ingest(shoeprint_maps),
                pairs,
                rotations,
                scales,
                precision=precision,
                k=top_k,
            )
        if rank == 0 and len(ranks):
            # the lines similarity.py:375 prints, same text and order, in ONE write: tqdm clears and redraws its bars around
            # every write call (~60 us each, 90 ms for 1,500 shoemarks)
            pbar.write("\n".join(f"Print {shoemark_id} true match ranked {rk}" for shoemark_id, rk in enumerate(ranks)))
        pbar.update(pbar.total)
    last_result.clear()
    last_result.update(scores=scores, topk=topk, h2d_bytes=engine.last_h2d_bytes)
    return ranks


def _ncc_surfaces(marks: np.ndarray, prints: np.ndarray) -> np.ndarray:
    """Per-channel NCC surfaces ``[C, Hp, Wp]`` (float32) of UNCROPPED maps ``[C,h,w]`` through the library's pack
    kernels and ``sir_ncc_surface`` (fp32 CUDA cores, one small launch per channel).  Helper path only: the matching
    path never materialises a surface."""
    import ctypes as C

    import torch

    from . import _native as nat

    c, h, w = marks.shape
    hm, wm = h - 2 * _PAD, w - 2 * _PAD
    gal = engine.MapSet.from_host([prints])
    ops = engine.GalleryOperands.pack(gal.groups[0], keep_fp32=True)
    rn = ops.rnorm(hm, wm, simt=True)
    dev = rn.device
    tmap = torch.from_numpy(np.ascontiguousarray(marks)[None]).to(dev)
    kpad = int(nat.lib.sir_template_kpad(hm, wm))
    thi = torch.empty((c, 1, kpad), dtype=torch.float16, device=dev)
    tlo = torch.empty_like(thi)
    t32 = torch.empty((c, 1, hm * wm), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.sir_template_pack(C.c_void_p(tmap.data_ptr()), 1, c, h, w, 0, 1, C.c_void_p(thi.data_ptr()),
                                        C.c_void_p(tlo.data_ptr()), C.c_void_p(t32.data_ptr()), st), "sir_template_pack")
    m = ops.Hp * ops.Wp
    out = torch.empty((c, ops.Hp, ops.Wp), dtype=torch.float32, device=dev)
    for ch in range(c):
        nat.check(nat.lib.sir_ncc_surface(C.c_void_p(ops.gz.data_ptr() + 4 * ch * m), C.c_void_p(rn.data_ptr() + 4 * ch * m), ops.Hp, ops.Wp,
                                          C.c_void_p(t32.data_ptr() + 4 * ch * hm * wm), hm, wm, C.c_void_p(out.data_ptr() + 4 * ch * m), st),
                  "sir_ncc_surface")
    return out.cpu().numpy()


def get_similarity(shoemark: FeatureMapsArrayType, shoeprint: FeatureMapsArrayType) -> np.floating:
    """Similarity of one shoemark against one shoeprint (``similarity.py:75-108``): per-channel NCC
    of the 2-cell-cropped maps, summed over channels, max over positions, divided by C.

    The fused path floors at 0 like ``compare_maps`` does (``similarity.py:355``); a lone ``get_similarity`` call in the
    reference does not, so a pair whose best position is not positive is re-evaluated from its per-channel surfaces and
    the (negative) maximum is returned as the reference would."""
    mark, prnt = _as_f32(shoemark), _as_f32(shoeprint)
    probes = engine.MapSet.from_host([mark])
    gallery = engine.MapSet.from_host([prnt])
    score = float(engine.score_matrix(probes, gallery, None, None, engine.DEFAULT_PRECISION)[0, 0].item())
    if score > 0.0:
        return np.float64(score)
    surf = _ncc_surfaces(mark, prnt).astype(np.float64).sum(axis=0)
    return np.float64(surf.max() / mark.shape[0])


def normxcorr(
    template: ImageArrayType,
    image: ImageArrayType,
    mode: Literal["full", "valid", "same"] = "same",
) -> ImageArrayType:
    """Normalised cross-correlation surface of ``template`` over ``image`` (``similarity.py:26-72``).

    Helper, not on the matching path: the fused kernel never materialises the surface
    (``compare_maps`` / ``get_similarity`` go through the fused kernels).  Here the zero-meaned
    operands and the window norm come from the library's pack kernels and the surface from
    ``sir_ncc_surface`` (fp32 CUDA cores).  Only ``mode="same"`` -- the only mode the reference uses
    (``similarity.py:104``) -- is supported.
    """
    if mode != "same":
        raise NotImplementedError("only mode='same' is used by the matching path (similarity.py:104)")
    t = _as_f32(template)
    g = _as_f32(image)
    # frame both so the library's 2-cell crop removes exactly the frame
    return _ncc_surfaces(np.pad(t, _PAD)[None], np.pad(g, _PAD)[None])[0].astype(np.float64)


def _get_rank(similarities, matching_pairs: list[int], print_id: int) -> int:
    """1-based rank of the true match in one score row (``similarity.py:378-386``)."""
    row = np.asarray(similarities)
    return int((row > row[matching_pairs[print_id]]).sum()) + 1
