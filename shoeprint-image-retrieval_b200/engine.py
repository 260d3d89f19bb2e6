"""Device-side orchestration of the matching path (PyTorch owns memory and streams; every
kernel is in ``libsir.so``).

Flow of the default precision mode (reference call sites in brackets):

    gallery maps --sir_gallery_pack_f32--> fp16 operands + float32 copy        [similarity.py:92,49]
    probe maps --sir_variant_resize (scaled variants only)--> variant maps     [similarity.py:262-276, 321-353]
               --sir_template_pack_screen (rotation / transposition through an index map)--> column blocks
                                                                               [similarity.py:267,92,48,67]
    per (template shape, gallery shape): sir_gallery_window_rnorm              [similarity.py:57-65]
        sir_ncc_screen (fp16 tensor-core surface, max-fused, candidate records) [similarity.py:53-55,68,100-108,355-367]
        sir_ncc_refine (exact float32 value at the candidate positions)
    scores --sir_true_scores / sir_rank_topk--> ranks, top-k                   [similarity.py:378-386]

The single-pass modes (``fp16_fp8c``, ``fp16x3``, ``fp16x1``, ``fp32_simt``) materialise every variant
(sir_variant_rotate / resize), pack with sir_template_pack[_fp8c] and score with sir_ncc_scores[_fp8c].

Ragged inputs are grouped by shape on the host (the reference never pads: dataloader.py:231-237).
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _native as nat

__all__ = [
    "MapGroup",
    "MapSet",
    "GalleryOperands",
    "variant_plan",
    "scaled_size",
    "score_matrix",
    "rank_true_matches",
    "compare",
    "launch_counter",
]

EDGE = 2  # similarity.py:92-93

#: parity-grade default: screening on the tensor cores with plain fp16 operands (one MMA per K step), then the
#: positions that can hold a pair's maximum are re-evaluated exactly in float32 (``sir_ncc_screen`` + ``sir_ncc_refine``;
#: <= 2e-6 relative vs the float64 oracle).  "fp16_fp8c" (fp16 + fp8 correction MMAs, 2 MMA-equivalents per MAC) and
#: "fp16x3" (3) are the single-pass parity-grade modes, "fp16x1" the screening pass alone (~2e-4), "fp32_simt" the
#: CUDA-core check path.
DEFAULT_PRECISION = "fp16_refine"

#: candidate margin of the screening pass, tau(m) = TAU_REL*|m| + TAU_ABS in score units: a position is re-evaluated
#: when its screened value is within tau of the pair's screened maximum.  The screening error is <= ~2e-4 relative
#: (measured, tests/precision_study.py) and <= ~1e-5 absolute on near-zero scores; tau covers twice that with margin.
TAU_REL = 1.0e-3
TAU_ABS = 2.0e-5

#: probes uploaded from host lists travel in chunks of this many maps; a chunk's columns are launched once at least
#: FLUSH_MIN_COLS of one template shape have accumulated, while the next chunk is still being copied
PROBE_CHUNK = int(os.environ.get("SIR_PROBE_CHUNK", "512"))
FLUSH_MIN_COLS = int(os.environ.get("SIR_FLUSH_MIN_COLS", "1024"))


class _LaunchCounter:
    """Counts libsir kernel launches (bench.py reports it as ``gpu_launches``)."""

    def __init__(self) -> None:
        self.n = 0

    def add(self, k: int = 1) -> None:
        self.n += k


launch_counter = _LaunchCounter()

#: when set to a list, every sir_ncc_scores / sir_ncc_screen launch appends (start_event, end_event, algorithmic_flops)
#: recorded on the launching stream (bench.py uses this for the roofline line); refine_events likewise collects
#: (start_event, end_event) of every sir_ncc_refine launch
kernel_events: list | None = None
refine_events: list | None = None
#: device counters of sir_ncc_refine ([0] positions evaluated, [1] records listing > 3 rows, [2] tiles with work), created on
#: first use when ``collect_refine_stats`` is set
collect_refine_stats = False
_refine_stats: dict = {}


def refine_stats(reset: bool = False) -> dict:
    """Counters accumulated by sir_ncc_refine on the current device since the last reset."""
    dev = torch.cuda.current_device()
    t = _refine_stats.get(dev)
    if t is None:
        return {"positions": 0, "dense_records": 0, "tiles": 0}
    v = t.cpu().tolist()
    if reset:
        t.zero_()
    return {"positions": v[0], "dense_records": v[1], "tiles": v[2]}


def _stats_ptr() -> C.c_void_p:
    if not collect_refine_stats:
        return C.c_void_p(0)
    dev = torch.cuda.current_device()
    if dev not in _refine_stats:
        _refine_stats[dev] = torch.zeros(4, dtype=torch.int64, device="cuda")
    return C.c_void_p(_refine_stats[dev].data_ptr())


def _zeros(shape, dtype: torch.dtype, dev) -> torch.Tensor:
    """Zero-filled device tensor through ``sir_memset_zero`` (cudaMemsetAsync on the current stream)."""
    t = torch.empty(shape, dtype=dtype, device=dev)
    nat.check(nat.lib.sir_memset_zero(_ptr(t), t.numel() * t.element_size(), _stream()), "sir_memset_zero")
    return t


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_current_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> C.c_void_p:
    """The current CUDA stream as a ``cudaStream_t``.  Ragged passes make ~17,000 library calls; ``torch.cuda.current_stream()``
    builds a Python Stream object (14 us) per call, the raw accessor is ~0.3 us."""
    if _raw_stream is not None and _current_device is not None:
        return C.c_void_p(_raw_stream(_current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("the matching path needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# --------------------------------------------------------------------------- inputs

class _Pending:
    """An upload in flight: ``wait()`` blocks the host until the copy has been queued and makes the current stream
    wait for it."""

    def __init__(self) -> None:
        self.flag = threading.Event()
        self.cuda_event: torch.cuda.Event | None = None
        self.error: BaseException | None = None

    def wait(self) -> None:
        self.flag.wait()
        if self.error is not None:
            raise self.error
        torch.cuda.current_stream().wait_event(self.cuda_event)


class _Uploader:
    """Host -> device copies of feature-map lists from ordinary (pageable) numpy arrays.

    ``cudaMemcpy`` from pageable memory is staged by the driver on one thread (~10 GB/s) and blocks the caller.  Here a
    background thread fills a ring of page-locked buffers with several copy threads (numpy releases the GIL) and queues
    one large asynchronous DMA per buffer on a side stream, so the caller goes on launching kernels for the maps that
    have already arrived (``engine.compare`` scores probe chunk i while chunk i+1 is on its way)."""

    BUF_BYTES = 64 << 20
    NBUF = 3
    # copy threads per process: half the host cores, shared among the ranks of the box (torchrun sets LOCAL_WORLD_SIZE)
    THREADS = max(2, min(8, (os.cpu_count() or 2) // (2 * max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))))
    _instances: dict = {}

    @classmethod
    def get(cls, dev: torch.device) -> "_Uploader":
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in cls._instances:
            cls._instances[key] = cls(torch.device("cuda", key))
        return cls._instances[key]

    def __init__(self, dev: torch.device) -> None:
        self.dev = dev
        self.stream = torch.cuda.Stream(device=dev)
        self.bufs = [torch.empty(self.BUF_BYTES, dtype=torch.uint8).pin_memory() for _ in range(self.NBUF)]
        self.free_ev: list = [None] * self.NBUF
        self.next = 0
        self.copy_pool = ThreadPoolExecutor(self.THREADS)
        self.worker = ThreadPoolExecutor(1)

    def submit(self, arrays: list[np.ndarray], dst: torch.Tensor, after: torch.cuda.Event) -> _Pending:
        pend = _Pending()
        dst.record_stream(self.stream)
        self.worker.submit(self._run, arrays, dst, after, pend)
        return pend

    def _run(self, arrays: list[np.ndarray], dst: torch.Tensor, after: torch.cuda.Event, pend: _Pending) -> None:
        try:
            torch.cuda.set_device(self.dev)
            self.stream.wait_event(after)  # the destination may be memory the caller's stream has only just released
            shape = tuple(arrays[0].shape)
            per = int(np.prod(shape)) * 4
            with torch.cuda.stream(self.stream):
                if per > self.BUF_BYTES:  # a single map larger than a staging buffer: plain copies
                    for j, a in enumerate(arrays):
                        dst[j].copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
                else:
                    n_buf = self.BUF_BYTES // per
                    for j0 in range(0, len(arrays), n_buf):
                        b = self.next
                        self.next = (b + 1) % self.NBUF
                        if self.free_ev[b] is not None:
                            self.free_ev[b].synchronize()
                        chunk = arrays[j0 : j0 + n_buf]
                        view = self.bufs[b][: len(chunk) * per].view(torch.float32).view(len(chunk), *shape)
                        host = view.numpy()

                        def fill(lo: int, hi: int, host=host, chunk=chunk) -> None:
                            for k in range(lo, hi):
                                np.copyto(host[k], chunk[k], casting="same_kind")

                        parts = min(self.THREADS, len(chunk))
                        jobs = [self.copy_pool.submit(fill, len(chunk) * t // parts, len(chunk) * (t + 1) // parts) for t in range(parts)]
                        for job in jobs:
                            job.result()
                        dst[j0 : j0 + len(chunk)].copy_(view, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.stream)
                        self.free_ev[b] = ev
                done = torch.cuda.Event()
                done.record(self.stream)
            pend.cuda_event = done
        except BaseException as exc:  # noqa: BLE001 - re-raised in the caller's thread by wait()
            pend.error = exc
        finally:
            pend.flag.set()


@dataclass
class MapGroup:
    """Feature maps of one shape: ``maps [n, C, h, w]`` float32 on the device, ``ids`` = their
    positions in the caller's list.  ``pending``: an upload still in flight (``wait()`` before the maps are read)."""

    maps: torch.Tensor
    ids: torch.Tensor  # int64 [n], host
    pending: _Pending | None = None

    @property
    def shape_hw(self) -> tuple[int, int]:
        return int(self.maps.shape[2]), int(self.maps.shape[3])

    def wait(self) -> None:
        if self.pending is not None:
            self.pending.wait()
            self.pending = None


def transpose_maps(maps: torch.Tensor) -> torch.Tensor:
    """``[n,C,h,w] -> [n,C,w,h]`` through ``sir_maps_transpose``."""
    n, c, h, w = (int(v) for v in maps.shape)
    out = torch.empty((n, c, w, h), dtype=torch.float32, device=maps.device)
    nat.check(nat.lib.sir_maps_transpose(_ptr(maps), n, c, h, w, _ptr(out), _stream()), "sir_maps_transpose")
    launch_counter.add()
    return out


def plan_cost(precision: int, g: int, hp: int, wp: int, hm: int, wm: int) -> float:
    """Estimated SM cycles per (gallery, 256-column tile, channel) from the library's own planner
    (``sir_ncc_cost``); ``inf`` when the shape does not fit that mode's shared-memory plan."""
    cost = C.c_double(0.0)
    rc = nat.lib.sir_ncc_cost(precision, g, hp, wp, hm, wm, C.byref(cost))
    return cost.value if rc == 0 else float("inf")


@dataclass
class MapSet:
    groups: list[MapGroup]
    count: int
    channels: int
    h2d_bytes: int = 0

    @staticmethod
    def from_host(maps: list[np.ndarray], chunk: int | None = None, lazy: bool = False) -> "MapSet":
        """Group a list of ``[C,h,w]`` float32 arrays by shape and upload them (``_Uploader``: pinned staging ring filled
        by several threads, asynchronous DMA on a side stream).  ``chunk``: cut shape groups into pieces of at most that
        many maps, each its own ``MapGroup``, so that the first piece can be used while the rest is still on its way;
        ``lazy``: return before the copies have landed (every group carries its ``pending`` handle and must be
        ``wait()``-ed for before its maps are read)."""
        dev = _require_cuda()
        if len(maps) == 0:
            raise ValueError("empty list of feature maps")
        resident = maps.device_copies() if hasattr(maps, "device_copies") else None
        if resident is not None:  # the feature stage left these maps on the device (network.FeatureMapList): no upload
            groups = []
            for tensor, idx in resident:
                shp = tuple(int(v) for v in tensor.shape[1:])
                if len(shp) != 3 or shp[1] <= 2 * EDGE or shp[2] <= 2 * EDGE:
                    raise ValueError(f"feature map of shape {shp} vanishes after the 2-cell crop (similarity.py:92-93)")
                groups.append(MapGroup(tensor.contiguous(), torch.tensor(idx, dtype=torch.int64)))
            chans = {int(g.maps.shape[1]) for g in groups}
            if len(chans) != 1:
                raise ValueError(f"feature maps disagree on the channel count: {sorted(chans)}")
            return MapSet(groups, len(maps), chans.pop(), 0)
        by_shape: dict[tuple[int, int, int], list[int]] = {}
        for i, m in enumerate(maps):
            if m.ndim != 3:
                raise ValueError(f"feature map {i} has shape {m.shape}, expected [C,h,w]")
            by_shape.setdefault(tuple(m.shape), []).append(i)
        chans = {s[0] for s in by_shape}
        if len(chans) != 1:
            raise ValueError(f"feature maps disagree on the channel count: {sorted(chans)}")
        for shp in by_shape:
            if shp[1] <= 2 * EDGE or shp[2] <= 2 * EDGE:
                raise ValueError(f"feature map of shape {shp} vanishes after the 2-cell crop (similarity.py:92-93)")
        # every destination is allocated before the first copy is queued: the side stream then only has to wait for what
        # the caller's stream had queued up to here
        pieces = []
        for shp, idx in by_shape.items():
            # the first pieces are smaller (chunk/4, chunk/2): the consumer's first launch waits for the first piece only
            s0, step = 0, (len(idx) if not chunk else max(1, chunk // 4))
            while s0 < len(idx):
                part = idx[s0 : s0 + step]
                pieces.append((part, torch.empty((len(part), *shp), dtype=torch.float32, device=dev)))
                s0 += step
                if chunk:
                    step = min(max(1, chunk), 2 * step)
        after = torch.cuda.Event()
        after.record()
        up = _Uploader.get(dev)
        groups, nbytes = [], 0
        for part, dst in pieces:
            arrays = []
            for i in part:
                src = maps[i]
                if src.dtype != np.float32:
                    src = np.asarray(src, dtype=np.float32)
                arrays.append(src)
            groups.append(MapGroup(dst, torch.tensor(part, dtype=torch.int64), up.submit(arrays, dst, after)))
            nbytes += dst.numel() * 4
        if not lazy:
            for g in groups:
                g.wait()
        return MapSet(groups, len(maps), chans.pop(), nbytes)

    @staticmethod
    def from_device(maps: torch.Tensor) -> "MapSet":
        """One uniform-shape group already resident on the device: ``[n, C, h, w]`` float32."""
        if maps.dim() != 4 or maps.dtype != torch.float32 or not maps.is_cuda:
            raise ValueError("expected a CUDA float32 tensor [n,C,h,w]")
        maps = maps.contiguous()
        return MapSet([MapGroup(maps, torch.arange(maps.shape[0]))], int(maps.shape[0]), int(maps.shape[1]))


@dataclass
class GalleryOperands:
    """Packed gallery group (K5) + cache of window inverse norms per template shape (K6)."""

    G: int
    C: int
    Hp: int
    Wp: int
    ghi: torch.Tensor
    glo: torch.Tensor
    gexp: torch.Tensor
    gz: torch.Tensor | None
    ids: torch.Tensor
    g32: torch.Tensor | None = None       # float32 copy of the scaled operand (exact re-evaluation of fp16_refine)
    _rnorm: dict = field(default_factory=dict)
    _fp8: tuple | None = None
    _source: MapGroup | None = None       # kept so the other orientation can be packed on demand
    _keep_fp32: bool = False
    _transposed: "GalleryOperands | None" = None

    def transposed(self) -> "GalleryOperands":
        """The same gallery group packed from transposed maps (built on first use)."""
        if self._transposed is None:
            if self._source is None:
                raise RuntimeError("this gallery pack does not hold its source maps")
            grp = MapGroup(transpose_maps(self._source.maps), self._source.ids)
            self._transposed = GalleryOperands.pack(grp, self._keep_fp32, self.g32 is not None)
            self._transposed._source = None
        return self._transposed

    def fp8_companions(self) -> tuple[torch.Tensor, torch.Tensor]:
        """e4m3 copies (hi/4, lo*4) of the packed gallery for the fp8-corrected mode, built once."""
        if self._fp8 is None:
            wp8 = int(nat.lib.sir_gallery_pitch8(self.Wp))
            g8a = torch.empty((self.G, self.C, self.Hp, wp8), dtype=torch.uint8, device=self.ghi.device)
            g8l = torch.empty_like(g8a)
            nat.check(
                nat.lib.sir_gallery_pack_fp8c(_ptr(self.ghi), _ptr(self.glo), self.G, self.C, self.Hp, self.Wp, _ptr(g8a), _ptr(g8l), _stream()),
                "sir_gallery_pack_fp8c",
            )
            launch_counter.add()
            self._fp8 = (g8a, g8l)
        return self._fp8

    @staticmethod
    def pack(group: MapGroup, keep_fp32: bool, with_f32: bool = False) -> "GalleryOperands":
        n, c, h, w = (int(v) for v in group.maps.shape)
        hp, wp = h - 2 * EDGE, w - 2 * EDGE
        dev = group.maps.device
        ghi = torch.empty((n, c, hp, int(nat.lib.sir_gallery_pitch(wp))), dtype=torch.float16, device=dev)
        glo = torch.empty_like(ghi)
        gexp = torch.empty((n, c), dtype=torch.int32, device=dev)
        gz = torch.empty((n, c, hp, wp), dtype=torch.float32, device=dev) if keep_fp32 else None
        g32 = torch.empty(ghi.shape, dtype=torch.float32, device=dev) if with_f32 else None
        nat.check(
            nat.lib.sir_gallery_pack_f32(_ptr(group.maps), n, c, h, w, _ptr(ghi), _ptr(glo), _ptr(gexp), _ptr(gz), _ptr(g32), _stream()),
            "sir_gallery_pack_f32",
        )
        launch_counter.add()
        ops = GalleryOperands(n, c, hp, wp, ghi, glo, gexp, gz, group.ids, g32)
        ops._source, ops._keep_fp32 = group, keep_fp32
        return ops

    def rnorm(self, hm: int, wm: int, simt: bool) -> torch.Tensor:
        key = (hm, wm, simt)
        if key not in self._rnorm:
            if len(self._rnorm) >= 2:  # a window table is as large as the gallery: keep few
                self._rnorm.pop(next(iter(self._rnorm)))
            out = torch.empty((self.G, self.C, self.Hp * self.Wp), dtype=torch.float32, device=self.ghi.device)
            nat.check(
                nat.lib.sir_gallery_window_rnorm(
                    _ptr(self.ghi), _ptr(self.glo), _ptr(self.gz if simt else None),
                    self.G, self.C, self.Hp, self.Wp, hm, wm, _ptr(out), _stream(),
                ),
                "sir_gallery_window_rnorm",
            )
            launch_counter.add()
            self._rnorm[key] = out
        return self._rnorm[key]


# --------------------------------------------------------------------------- variants

def variant_plan(rotations, scales) -> list[tuple[float | None, float | None]]:
    """(rotation, scale) of every variant the reference scores (similarity.py:321-353 with the
    list handling of :282): with both given, ``[id] + [scale_s(v) for v in (id, rot...) for s]``,
    i.e. the rotated-only variants are not scored."""
    if rotations is None and scales is None:
        return [(None, None)]
    if scales is None:
        return [(None, None)] + [(float(r), None) for r in rotations]
    if rotations is None:
        return [(None, None)] + [(None, float(s)) for s in scales]
    plan: list[tuple[float | None, float | None]] = [(None, None)]
    for r in [None, *rotations]:
        plan.extend((None if r is None else float(r), float(s)) for s in scales)
    return plan


def scaled_size(h: int, w: int, s: float) -> tuple[int, int]:
    """Target size of ``Image.resize((int(w*s), int(h*s)))`` (similarity.py:269-274)."""
    return int(h * s), int(w * s)


def make_variant(maps: torch.Tensor, rot: float | None, scale: float | None) -> torch.Tensor:
    """Rotate (nearest, Pillow fixed point) then resize (bicubic) a group ``[n,C,h,w]``."""
    n, c, h, w = (int(v) for v in maps.shape)
    out = maps
    if rot is not None:
        rotated = torch.empty_like(maps)
        nat.check(nat.lib.sir_variant_rotate(_ptr(out), n, c, h, w, float(rot), _ptr(rotated), _stream()), "sir_variant_rotate")
        launch_counter.add()
        out = rotated
    if scale is not None:
        h2, w2 = scaled_size(h, w, scale)
        if h2 < 1 or w2 < 1:
            raise ValueError(f"scale {scale} shrinks a {h}x{w} map to nothing")
        if (h2, w2) != (h, w):
            resized = torch.empty((n, c, h2, w2), dtype=torch.float32, device=maps.device)
            two_pass = h2 != h and w2 != w
            tmp = torch.empty((n, c, h, w2), dtype=torch.float32, device=maps.device) if two_pass else None
            ws_bytes = int(nat.lib.sir_variant_resize_workspace_bytes(h, w, h2, w2))
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=maps.device)  # tap tables of the passes
            nat.check(
                nat.lib.sir_variant_resize(_ptr(out), n, c, h, w, h2, w2, _ptr(resized), _ptr(tmp), _ptr(ws), ws_bytes, _stream()),
                "sir_variant_resize",
            )
            launch_counter.add(2 if two_pass else 1)
            out = resized
    return out


def resize_images_lanczos(images: list[np.ndarray], sizes: list[tuple[int, int]]) -> list[np.ndarray]:
    """``np.array(Image.fromarray(im).resize((w2, h2), LANCZOS))`` for uint8 images on the device (``sir_image_resize_lanczos``,
    bit exact with Pillow's 8-bit resampler; ``dataloader.py:231-237``).  ``sizes[i] = (h2, w2)``.  Images that share source
    and target size travel as one batch."""
    dev = _require_cuda()
    out: list = [None] * len(images)
    batches: dict[tuple, list[int]] = {}
    for i, (im, hw) in enumerate(zip(images, sizes)):
        batches.setdefault((tuple(im.shape), tuple(hw)), []).append(i)
    for (shape, (h2, w2)), idx in batches.items():
        h, w = shape[:2]
        ch = 1 if len(shape) == 2 else int(shape[2])
        src = torch.from_numpy(np.stack([np.ascontiguousarray(images[i]) for i in idx])).to(dev, non_blocking=True)
        dst = torch.empty((len(idx), h2, w2) + ((ch,) if len(shape) == 3 else ()), dtype=torch.uint8, device=dev)
        tmp = torch.empty((len(idx), h, w2, ch), dtype=torch.uint8, device=dev) if (h2 != h and w2 != w) else None
        ws_bytes = int(nat.lib.sir_image_resize_workspace_bytes(h, w, h2, w2))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        nat.check(nat.lib.sir_image_resize_lanczos(_ptr(src), len(idx), h, w, ch, h2, w2, _ptr(dst), _ptr(tmp), _ptr(ws), ws_bytes, _stream()),
                  "sir_image_resize_lanczos")
        launch_counter.add(2 if tmp is not None else 1)
        host = dst.cpu().numpy()
        for j, i in enumerate(idx):
            out[i] = host[j]
    return out


# --------------------------------------------------------------------------- scoring

@dataclass
class _Variant:
    """Columns of one (probe group, variant), generated on demand.

    ``source [n,C,h,w]`` are the probe maps; the variant is rotate(``rot``) then resize(``scale``) of them
    (similarity.py:267-274).  A rotation alone keeps the shape and is a pure gather (Pillow nearest), so with ``gather`` the
    screening mode lets the template pack apply it on the way in (``sir_variant_index_map``) and the rotated maps are never
    written: ``maps`` is then the source and ``rot`` the rotation still to apply.  Everything else is materialised the first
    time ``maps`` is read -- inside the block / bucket that scores it, so the thousands of small variant kernels of a ragged
    probe set are queued bucket by bucket behind each other's correlation launches instead of all up front."""

    source: torch.Tensor
    rot_angle: float | None = None
    scale: float | None = None
    gather: bool = False
    _made: torch.Tensor | None = None

    def __post_init__(self) -> None:
        n, c, h, w = (int(v) for v in self.source.shape)
        self.n = n
        self.shape_hw = (h, w)
        if self.scale is not None:
            self.shape_hw = scaled_size(h, w, self.scale)
            if self.shape_hw[0] < 1 or self.shape_hw[1] < 1:
                raise ValueError(f"scale {self.scale} shrinks a {h}x{w} map to nothing")
        self.gather = self.gather and self.scale is None

    @property
    def rot(self) -> float | None:
        """The rotation the template pack still has to apply (gather variants only)."""
        return self.rot_angle if self.gather else None

    @property
    def maps(self) -> torch.Tensor:
        if self.gather or (self.rot_angle is None and self.scale is None):
            return self.source
        if self._made is None:
            self._made = make_variant(self.source, self.rot_angle, self.scale)
        return self._made

    def materialised(self) -> torch.Tensor:
        """The variant as a tensor, whatever the mode (the single-pass modes pack from materialised maps)."""
        if self.gather and self.rot_angle is not None:
            if self._made is None:
                self._made = make_variant(self.source, self.rot_angle, None)
            return self._made
        return self.maps


_index_maps: dict = {}


def _gather_map(dev, h: int, w: int, rot: float | None, flip: bool) -> torch.Tensor | None:
    """Device index map of (rotation, transposition) for an ``h x w`` source map; None when it is the identity."""
    if rot is None and not flip:
        return None
    key = (dev.index, h, w, 0.0 if rot is None else float(rot), flip)
    if key not in _index_maps:
        if len(_index_maps) > 256:
            _index_maps.clear()
        m = torch.empty(h * w, dtype=torch.int32, device=dev)
        nat.check(nat.lib.sir_variant_index_map(h, w, key[3], 1 if flip else 0, _ptr(m), _stream()), "sir_variant_index_map")
        launch_counter.add()
        _index_maps[key] = m
    return _index_maps[key]


@dataclass
class _Block:
    """Pending columns of one template shape."""

    maps: list[_Variant] = field(default_factory=list)
    ids: list[torch.Tensor] = field(default_factory=list)
    ncols: int = 0


def _score_block(block: _Block, hw: tuple[int, int], gallery: list[GalleryOperands], offsets: list[int],
                 scores: torch.Tensor, precision: int, approx: torch.Tensor | None = None) -> None:
    """Scores one column block against every gallery group.

    Scores are invariant under transposing probe and gallery maps alike, and ``fp16_fp8c`` and
    ``fp16x3`` are both parity grade, so each (block, gallery group) runs in the cheapest of those
    configurations according to the library's planner: the kernel tiles positions 16 x 8 and template
    rows by 8 / 16 taps, and wide templates make the fp8-corrected mode generator bound."""
    h, w = hw
    hm, wm = h - 2 * EDGE, w - 2 * EDGE
    plans: dict[tuple[int, bool], list[tuple[GalleryOperands, int]]] = {}
    for ops, g0 in zip(gallery, offsets):
        if precision == nat.PREC_FP32_SIMT:
            plans.setdefault((precision, False), []).append((ops, g0))
            continue
        modes = [precision, nat.PREC_FP16X3] if precision == nat.PREC_FP16_FP8C else [precision]
        best, best_cost = (modes[-1], False), float("inf")
        for mode in modes:
            for flip in ((False, True) if ops._source is not None else (False,)):
                cost = plan_cost(mode, ops.G, *((ops.Wp, ops.Hp, wm, hm) if flip else (ops.Hp, ops.Wp, hm, wm)))
                if cost < best_cost * (0.97 if (mode, flip) != (modes[0], False) else 1.0):
                    best, best_cost = (mode, flip), cost
        if best_cost == float("inf"):
            raise nat.SirError(f"template {hm}x{wm} does not fit any shared-memory plan of the correlation kernel")
        plans.setdefault(best, []).append((ops, g0))
    for (mode, flip), members in plans.items():
        ops = [o.transposed() if flip else o for o, _ in members]
        _score_block_oriented(block, (h, w), flip, ops, [g for _, g in members], scores, mode, approx)


def _score_block_oriented(block: _Block, hw: tuple[int, int], flip: bool, gallery: list[GalleryOperands], offsets: list[int],
                          scores: torch.Tensor, precision: int, approx: torch.Tensor | None = None) -> None:
    """``hw``: shape of the block's variants as the caller holds them; ``flip``: the launch works on transposed maps
    (``gallery`` already is).  The screening mode packs straight from the source maps through an index map; the other
    modes materialise the rotated / transposed variant first."""
    h0, w0 = hw
    h, w = (w0, h0) if flip else (h0, w0)
    hm, wm = h - 2 * EDGE, w - 2 * EDGE
    dev = scores.device
    c = gallery[0].C
    ncols = block.ncols
    simt = precision == nat.PREC_FP32_SIMT
    fp8c = precision == nat.PREC_FP16_FP8C
    refine = precision == nat.PREC_FP16_REFINE
    if refine and approx is None:
        raise ValueError("fp16_refine needs the approximate score buffer")
    kpad = int(nat.lib.sir_template_kpad_fp8c(hm, wm) if fp8c else nat.lib.sir_template_kpad(hm, wm))
    thi = torch.empty((c, ncols, kpad), dtype=torch.float16, device=dev)
    tlo = None if (fp8c or refine) else torch.empty_like(thi)
    t32p = torch.empty((c, ncols, kpad), dtype=torch.float32, device=dev) if refine else None
    t8b = torch.empty((c, ncols, kpad), dtype=torch.uint8, device=dev) if fp8c else None
    t8l = torch.empty_like(t8b) if fp8c else None
    t32 = torch.empty((c, ncols, hm * wm), dtype=torch.float32, device=dev) if simt else None
    col0 = 0
    for var in block.maps:
        n = var.n
        if refine:
            nat.check(nat.lib.sir_template_pack_screen(_ptr(var.maps), n, c, h, w, hm, wm, col0, ncols, _ptr(thi), _ptr(t32p),
                                                       _ptr(_gather_map(dev, h0, w0, var.rot, flip)), _stream()), "sir_template_pack_screen")
            launch_counter.add()
            col0 += n
            continue
        m = var.materialised()
        if flip:
            m = transpose_maps(m)
        if fp8c:
            nat.check(
                nat.lib.sir_template_pack_fp8c(_ptr(m), n, c, h, w, col0, ncols, _ptr(thi), _ptr(t8b), _ptr(t8l), _stream()),
                "sir_template_pack_fp8c",
            )
        else:
            nat.check(
                nat.lib.sir_template_pack(_ptr(m), n, c, h, w, col0, ncols, _ptr(thi), _ptr(tlo), _ptr(t32), _stream()),
                "sir_template_pack",
            )
        launch_counter.add()
        col0 += n
    ids_all = torch.cat(block.ids)
    variants_hint = max(1, int(ids_all.numel()) // max(1, int(ids_all.unique().numel())))  # columns per probe in this block
    col2probe = ids_all.to(torch.int32).to(dev, non_blocking=True)
    for ops, g0 in zip(gallery, offsets):
        rn = ops.rnorm(hm, wm, simt)
        if kernel_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if refine:
            rec = torch.empty((int(nat.lib.sir_ncc_screen_rec_count(ops.G, ops.Hp, ops.Wp, ncols)), 2), dtype=torch.int32, device=dev)
            nat.check(
                nat.lib.sir_ncc_screen(
                    _ptr(ops.ghi), _ptr(rn), None, ops.G, ops.C, ops.Hp, ops.Wp, _ptr(thi), ncols, ncols, hm, wm,
                    _ptr(col2probe), _ptr(approx), int(approx.stride(0)), g0, TAU_REL, TAU_ABS, _ptr(rec), None, _stream(),
                ),
                "sir_ncc_screen",
            )
        elif fp8c:
            g8a, g8l = ops.fp8_companions()
            nat.check(
                nat.lib.sir_ncc_scores_fp8c(
                    _ptr(ops.ghi), _ptr(g8a), _ptr(g8l), _ptr(rn), ops.G, ops.C, ops.Hp, ops.Wp,
                    _ptr(thi), _ptr(t8b), _ptr(t8l), ncols, ncols, hm, wm,
                    _ptr(col2probe), _ptr(scores), int(scores.stride(0)), g0, _stream(),
                ),
                "sir_ncc_scores_fp8c",
            )
        else:
            nat.check(
                nat.lib.sir_ncc_scores(
                    _ptr(ops.ghi), _ptr(ops.glo), _ptr(ops.gexp), _ptr(ops.gz), _ptr(rn),
                    ops.G, ops.C, ops.Hp, ops.Wp,
                    _ptr(thi), _ptr(tlo), _ptr(t32), ncols, ncols, hm, wm,
                    _ptr(col2probe), _ptr(scores), int(scores.stride(0)), g0, precision, _stream(),
                ),
                "sir_ncc_scores",
            )
        launch_counter.add()
        if kernel_events is not None:
            ev1.record()
            # 2 * C * (gallery positions) * (template taps) per (column, gallery)
            flops = 2.0 * ops.C * (ops.Hp * ops.Wp) * (hm * wm) * ncols * ops.G
            kernel_events.append((ev0, ev1, flops))
        if refine:
            if refine_events is not None:
                rv0, rv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                rv0.record()
            nat.check(
                nat.lib.sir_ncc_refine(
                    _ptr(ops.g32), _ptr(rn), None, ops.G, ops.C, ops.Hp, ops.Wp, _ptr(t32p), ncols, ncols, hm, wm,
                    _ptr(col2probe), _ptr(approx), _ptr(scores), int(scores.stride(0)), g0, TAU_REL, TAU_ABS, _ptr(rec), variants_hint,
                    _stats_ptr(), _stream(),
                ),
                "sir_ncc_refine",
            )
            launch_counter.add()
            if refine_events is not None:
                rv1.record()
                refine_events.append((rv0, rv1))


def _best_plan(ops: GalleryOperands, hm: int, wm: int, precision: int) -> tuple[int, bool, float]:
    """(mode, transposed, cost) of the cheapest parity-grade configuration for this shape (see _score_block)."""
    modes = [precision, nat.PREC_FP16X3] if precision == nat.PREC_FP16_FP8C else [precision]
    best, best_cost = (modes[-1], False), float("inf")
    for mode in modes:
        for flip in ((False, True) if ops._source is not None else (False,)):
            cost = plan_cost(mode, ops.G, *((ops.Wp, ops.Hp, wm, hm) if flip else (ops.Hp, ops.Wp, hm, wm)))
            if cost < best_cost * (0.97 if (mode, flip) != (modes[0], False) else 1.0):
                best, best_cost = (mode, flip), cost
    return best[0], best[1], best_cost


# Cost of one multi-shape launch, in the unit of ``plan_cost`` (SM cycles per gallery map, 256-column tile and channel),
# fitted on a B200 to 31 bucket launches of the configs[0] workload (``tools/ragged_buckets.py``): time = 0.62 us * (plan cost
# * column tiles + BUCKET_FIXED_COST) * G * C / (1175 * 80).  A partial tile still pays for the narrow MMA (~61 cycles whatever
# the width below 120 columns) and for the whole entry generation; the fixed part is the summed-area pass of the window
# norms, the refinement's start-up and the tail of the correlation kernel.
BUCKET_FIXED_COST = 29000.0
PARTIAL_TILE_FLOOR = 0.47


def _norm_chunk() -> int:
    return int(nat.lib.sir_ncc_norm_chunk())


def _pad_cols(n: int) -> int:
    """Columns of one true template shape inside a multi-shape block: whole window-norm chunks."""
    chunk = _norm_chunk()
    return -(-n // chunk) * chunk


_bucket_plans: dict = {}


def _tiles_equiv(cols: int) -> float:
    full, rest = divmod(cols, 256)
    return full + (max(PARTIAL_TILE_FLOOR, rest / 256) if rest else 0.0)


def _merge_buckets(buckets: dict[tuple, list], cost_of, max_cols: int, max_shapes: int) -> dict[tuple, list]:
    """Agglomerate the fine buckets ``(mode, flip, rows, cols) -> members``: two buckets of one mode and orientation are merged
    (bucket shape = the element-wise maximum, so the smaller templates carry more zero taps) whenever the model above says the
    merged launch is cheaper than the two -- few-column buckets cost almost as much as a full tile on their own."""
    def launch_cost(key, members) -> float:
        # the launch lays its columns out tallest template first and tells the kernel, per 256-column tile, which rows of the
        # bucket layout are occupied (d_tile_rows): a tile costs what its own tallest template costs
        mode, flip, bh, bw = key
        if cost_of(key) == float("inf"):  # the bucket shape itself has to fit the kernel's shared-memory plan
            return float("inf")
        heights = sorted(((w if flip else h) - 2 * EDGE, _pad_cols(blk.ncols)) for (h, w), blk in members)[::-1]
        cols = sum(n for _, n in heights)
        launches = max(1, -(-cols // max_cols), -(-len(members) // max_shapes))
        total, filled, tallest = 0.0, 0, 0
        for rows, n in heights:
            while n:
                take = min(n, 256 - filled)
                tallest = max(tallest, rows)
                filled += take
                n -= take
                if filled == 256:
                    total += cost_of((mode, flip, min(bh, -(-tallest // 8) * 8), bw))
                    filled, tallest = 0, 0
        if filled:
            total += cost_of((mode, flip, min(bh, -(-tallest // 8) * 8), bw)) * max(PARTIAL_TILE_FLOOR, filled / 256)
        return total + PARTIAL_TILE_FLOOR * (launches - 1) * cost_of(key) + launches * BUCKET_FIXED_COST

    items = {k: list(v) for k, v in buckets.items()}
    costs = {k: launch_cost(k, v) for k, v in items.items()}
    gains: dict[tuple, tuple] = {}  # (ka, kb) -> (gain, merged key, merged cost); entries die with either bucket

    def pair_gain(ka, kb):
        if (ka, kb) not in gains:
            km = (ka[0], ka[1], max(ka[2], kb[2]), max(ka[3], kb[3]))
            merged = launch_cost(km, items[ka] + items[kb] + (items[km] if km in items and km not in (ka, kb) else []))
            absorbed = costs[km] if km in items and km not in (ka, kb) else 0.0
            gains[(ka, kb)] = (costs[ka] + costs[kb] + absorbed - merged, km, merged)
        return gains[(ka, kb)]

    while len(items) > 1:
        best = None
        keys = list(items)
        for i, ka in enumerate(keys):
            for kb in keys[i + 1:]:
                if ka[:2] != kb[:2]:
                    continue
                gain, km, merged = pair_gain(ka, kb)
                if gain > 0 and (best is None or gain > best[0]):
                    best = (gain, ka, kb, km, merged)
        if best is None:
            break
        _, ka, kb, km, merged = best
        members = items.pop(ka) + items.pop(kb)
        costs.pop(ka), costs.pop(kb)
        if km in items:  # the merged shape is an existing bucket: it was part of the merged cost
            members = items.pop(km) + members
            costs.pop(km)
        items[km], costs[km] = members, merged
        for pair in [pr for pr in gains if ka in pr or kb in pr or km in pr]:
            del gains[pair]
    return items


def _score_buckets(blocks: dict[tuple[int, int], _Block], gallery: list[GalleryOperands], offsets: list[int],
                   scores: torch.Tensor, precision: int, max_cols: int = 8192, table_budget: int | None = None,
                   approx: torch.Tensor | None = None) -> None:
    """Multi-shape column tiles (``sir_template_pack_embed`` + ``sir_ncc_scores_multi``).

    Template shapes that round to the same bucket (rows to 8, columns to the mode's row alignment, in
    the orientation the planner prefers) are packed into one K layout, anchor on anchor; each shape's
    columns are padded to a multiple of ``sir_ncc_norm_chunk()`` (8) so that every such chunk has one true shape and
    hence one window-norm table.  Buckets are then merged while the launch-cost model gains (``_merge_buckets``)."""
    dev = scores.device
    if table_budget is None:  # window-norm tables of one launch: a quarter of what is free now, at most 24 GB
        table_budget = int(min(24 << 30, max(4 << 30, torch.cuda.mem_get_info(dev)[0] // 4)))
    for ops, g0 in zip(gallery, offsets):
        buckets: dict[tuple, list] = {}
        for (h, w), blk in blocks.items():
            hm, wm = h - 2 * EDGE, w - 2 * EDGE
            mode, flip, cost = _best_plan(ops, hm, wm, precision)
            if cost == float("inf"):
                raise nat.SirError(f"template {hm}x{wm} does not fit any shared-memory plan of the correlation kernel")
            oh, ow = (wm, hm) if flip else (hm, wm)
            align = 16 if mode == nat.PREC_FP16_FP8C else 8
            buckets.setdefault((mode, flip, -(-oh // 8) * 8, -(-ow // align) * align), []).append(((h, w), blk))
        table_bytes = ops.G * ops.C * ops.Hp * ops.Wp * 4
        max_shapes = max(1, table_budget // table_bytes)

        plan_costs: dict = {}

        def cost_of(key, ops=ops, memo=plan_costs) -> float:
            if key not in memo:
                mode, flip, bh, bw = key
                memo[key] = plan_cost(mode, ops.G, *((ops.Wp, ops.Hp) if flip else (ops.Hp, ops.Wp)), bh, bw)
            return memo[key]

        if os.environ.get("SIR_BUCKET_MERGE", "1") != "0":
            # the plan depends on the shapes and column counts only: a repeated call (next size cluster, next bench step) reuses it
            sig = (ops.G, ops.C, ops.Hp, ops.Wp, max_cols, max_shapes,
                   tuple(sorted((key, tuple(sorted((hw, blk.ncols) for hw, blk in members))) for key, members in buckets.items())))
            if sig not in _bucket_plans:
                if len(_bucket_plans) > 64:
                    _bucket_plans.clear()
                merged = _merge_buckets(buckets, cost_of, max_cols, max_shapes)
                _bucket_plans[sig] = {key: [hw for hw, _ in members] for key, members in merged.items()}
            by_shape = {hw: (hw, blk) for members in buckets.values() for hw, blk in members}
            buckets = {key: [by_shape[hw] for hw in shapes] for key, shapes in _bucket_plans[sig].items()}
        for (mode, flip, _, _), members in buckets.items():
            gops = ops.transposed() if flip else ops
            batch: list = []
            cols = 0
            for item in members:
                n32 = _pad_cols(item[1].ncols)
                if batch and (cols + n32 > max_cols or len(batch) + 1 > max_shapes):
                    _score_one_bucket(batch, gops, g0, scores, mode, flip, dev, approx)
                    batch, cols = [], 0
                batch.append(item)
                cols += n32
            if batch:
                _score_one_bucket(batch, gops, g0, scores, mode, flip, dev, approx)


def _score_one_bucket(members: list, gops: GalleryOperands, g0: int, scores: torch.Tensor, mode: int, flip: bool, dev,
                      approx: torch.Tensor | None = None) -> None:
    fp8c = mode == nat.PREC_FP16_FP8C
    refine = mode == nat.PREC_FP16_REFINE
    c = gops.C
    oriented = []
    for (h, w), blk in members:
        if refine:  # packed straight from the source maps through an index map (rotation, transposition)
            maps = [(v.maps, _gather_map(dev, h, w, v.rot, flip)) for v in blk.maps]
        else:
            maps = [(transpose_maps(v.materialised()) if flip else v.materialised(), None) for v in blk.maps]
        oriented.append(((w, h) if flip else (h, w), maps, blk))
    # tall templates first: a column tile then only spans the rows its own templates occupy (d_tile_rows below)
    oriented.sort(key=lambda item: (-item[0][0], -item[0][1]))
    hb = max(hw[0] for hw, _, _ in oriented) - 2 * EDGE
    wb = max(hw[1] for hw, _, _ in oriented) - 2 * EDGE
    ncols = sum(_pad_cols(blk.ncols) for _, _, blk in oriented)
    kpad = int(nat.lib.sir_template_kpad_fp8c(hb, wb) if fp8c else nat.lib.sir_template_kpad(hb, wb))
    thi = _zeros((c, ncols, kpad), torch.float16, dev)
    tlo = None if (fp8c or refine) else _zeros((c, ncols, kpad), torch.float16, dev)
    t32p = _zeros((c, ncols, kpad), torch.float32, dev) if refine else None
    t8b = _zeros((c, ncols, kpad), torch.uint8, dev) if fp8c else None
    t8l = _zeros((c, ncols, kpad), torch.uint8, dev) if fp8c else None
    col2probe = torch.zeros(ncols, dtype=torch.int32)
    ntiles = -(-ncols // 256)
    chunk = _norm_chunk()
    tab = torch.zeros(ntiles * (256 // chunk), dtype=torch.int64)
    tile_rows = torch.empty((ntiles, 2), dtype=torch.int32)
    tile_rows[:, 0], tile_rows[:, 1] = hb, 0
    # one pass over the gallery builds the window-norm tables of every member shape
    tables = [torch.empty((gops.G, gops.C, gops.Hp * gops.Wp), dtype=torch.float32, device=dev) for _ in oriented]
    ns = len(oriented)
    hms = (C.c_int * ns)(*[hw[0] - 2 * EDGE for hw, _, _ in oriented])
    wms = (C.c_int * ns)(*[hw[1] - 2 * EDGE for hw, _, _ in oriented])
    outs = (C.c_void_p * ns)(*[t.data_ptr() for t in tables])
    nat.check(nat.lib.sir_gallery_window_rnorm_multi(_ptr(gops.ghi), _ptr(gops.glo), gops.G, gops.C, gops.Hp, gops.Wp, ns, hms, wms, outs,
                                                     _stream()), "sir_gallery_window_rnorm_multi")
    launch_counter.add(-(-ns // 24))
    col0 = 0
    for ((h, w), maps, blk), rn in zip(oriented, tables):
        hm, wm = h - 2 * EDGE, w - 2 * EDGE
        start = col0
        for m, gmap in maps:
            n = int(m.shape[0])
            if refine:
                nat.check(nat.lib.sir_template_pack_screen(_ptr(m), n, c, h, w, hb, wb, col0, ncols, _ptr(thi), _ptr(t32p), _ptr(gmap), _stream()),
                          "sir_template_pack_screen")
            else:
                nat.check(nat.lib.sir_template_pack_embed(_ptr(m), n, c, h, w, hb, wb, col0, ncols, mode, _ptr(thi), _ptr(tlo), _ptr(t8b),
                                                          _ptr(t8l), _stream()), "sir_template_pack_embed")
            launch_counter.add()
            col0 += n
        col2probe[start:col0] = torch.cat(blk.ids).to(torch.int32)
        col0 = start + _pad_cols(blk.ncols)
        tab[start // chunk : col0 // chunk] = rn.data_ptr()
        # rows of the bucket layout this shape occupies (anchor on anchor: sir_pack.cu oy = Hb/2 - Hm/2), for every tile it touches
        t0, t1, oy = start // 256, (col0 - 1) // 256 + 1, hb // 2 - hm // 2
        tile_rows[t0:t1, 0] = torch.clamp(tile_rows[t0:t1, 0], max=oy)
        tile_rows[t0:t1, 1] = torch.clamp(tile_rows[t0:t1, 1], min=oy + hm)
    tab[col0 // chunk :] = tables[-1].data_ptr()
    d_tab = tab.to(dev, non_blocking=True)
    d_rows = tile_rows.to(dev, non_blocking=True) if os.environ.get("SIR_TILE_ROWS", "1") != "0" else None
    d_c2p = col2probe.to(dev, non_blocking=True)
    g8a, g8l = gops.fp8_companions() if fp8c else (None, None)
    if refine:
        rec = torch.empty((int(nat.lib.sir_ncc_screen_rec_count(gops.G, gops.Hp, gops.Wp, ncols)), 2), dtype=torch.int32, device=dev)
        nat.check(
            nat.lib.sir_ncc_screen(
                _ptr(gops.ghi), None, _ptr(d_tab), gops.G, gops.C, gops.Hp, gops.Wp, _ptr(thi), ncols, ncols, hb, wb,
                _ptr(d_c2p), _ptr(approx), int(approx.stride(0)), g0, TAU_REL, TAU_ABS, _ptr(rec), _ptr(d_rows), _stream(),
            ),
            "sir_ncc_screen",
        )
        nat.check(
            nat.lib.sir_ncc_refine(
                _ptr(gops.g32), None, _ptr(d_tab), gops.G, gops.C, gops.Hp, gops.Wp, _ptr(t32p), ncols, ncols, hb, wb,
                _ptr(d_c2p), _ptr(approx), _ptr(scores), int(scores.stride(0)), g0, TAU_REL, TAU_ABS, _ptr(rec),
                max(1, int(col2probe.numel()) // max(1, int(col2probe.unique().numel()))), _stats_ptr(), _stream(),
            ),
            "sir_ncc_refine",
        )
        launch_counter.add(2)
        return
    nat.check(
        nat.lib.sir_ncc_scores_multi(
            _ptr(gops.ghi), _ptr(None if fp8c else gops.glo), _ptr(g8a), _ptr(g8l), _ptr(d_tab), gops.G, gops.C, gops.Hp, gops.Wp,
            _ptr(thi), _ptr(tlo), _ptr(t8b), _ptr(t8l), ncols, ncols, hb, wb, _ptr(d_c2p), _ptr(scores), int(scores.stride(0)), g0,
            mode, _ptr(d_rows), _stream(),
        ),
        "sir_ncc_scores_multi",
    )
    launch_counter.add()


def score_matrix(probes: MapSet, gallery: MapSet, rotations=None, scales=None, precision: str = DEFAULT_PRECISION,
                 col_block: int = 16384, packed_gallery: list[GalleryOperands] | None = None,
                 gallery_chunk_bytes: int = 8 << 30, bucket_below: int = 192, operand_cache: dict | None = None) -> torch.Tensor:
    """float32 ``[Q, G]`` on the device: max over the variant set, floored at 0
    (``similarities_all`` of similarity.py:355-367), columns in the caller's gallery order.

    ``operand_cache``: a dict owned by the caller (``FeatureMapList.operand_cache``) that keeps the packed
    gallery operands of a gallery small enough to be packed up front, keyed by the operand kind."""
    dev = _require_cuda()
    if probes.channels != gallery.channels:
        raise ValueError(f"probe maps have {probes.channels} channels, gallery maps {gallery.channels}")
    prec = nat.PRECISIONS[precision]
    keep32 = prec == nat.PREC_FP32_SIMT
    with32 = prec == nat.PREC_FP16_REFINE
    # gallery groups are cut into chunks so that the per-chunk operands (packs, transposed copy, window
    # norms: ~3x the chunk's float32 bytes) stay bounded; small galleries are packed once up front,
    # large ones chunk by chunk inside every column block (packing is ~2 % of the correlation time)
    chunks: list[MapGroup] = []
    n_variants = len(variant_plan(rotations, scales))
    for grp in gallery.groups:
        grp.wait()
        n = int(grp.maps.shape[0])
        per_map = grp.maps[0].numel() * 4
        step = max(1, min(n, gallery_chunk_bytes // max(per_map, 1)))
        if with32:
            # the screening pass leaves 8 bytes per (column, gallery print, 16x8 patch): keep that record buffer under
            # ~16 GB for the widest column block this call can produce
            hp, wp = int(grp.maps.shape[2]) - 2 * EDGE, int(grp.maps.shape[3]) - 2 * EDGE
            patches = max(-(-hp // 16) * -(-wp // 8), -(-wp // 16) * -(-hp // 8))
            cols = min(max(col_block, 32768), max(1, probes.count * n_variants))
            step = max(1, min(step, (16 << 30) // (8 * patches * cols)))
        step = -(-n // -(-n // step))  # equal chunks
        for s0 in range(0, n, step):
            chunks.append(MapGroup(grp.maps[s0 : s0 + step], grp.ids[s0 : s0 + step]))
    offsets, g0 = [], 0
    for ch in chunks:
        offsets.append(g0)
        g0 += int(ch.maps.shape[0])
    total_bytes = sum(ch.maps.numel() * 4 for ch in chunks)
    prepacked = packed_gallery
    if prepacked is not None:
        # the caller's packs must be the packs of exactly these chunks, else columns would land in the wrong place
        if len(prepacked) != len(chunks) or any(o.G != int(ch.maps.shape[0]) or not torch.equal(o.ids, ch.ids) for o, ch in zip(prepacked, chunks)):
            raise ValueError("packed_gallery does not match the gallery's groups (it must hold one pack per shape group, "
                             f"each at most {gallery_chunk_bytes} bytes of maps)")
        if with32 and any(o.g32 is None for o in prepacked):
            raise ValueError("precision 'fp16_refine' needs gallery packs made with with_f32=True")
    elif total_bytes <= gallery_chunk_bytes:
        key = ("packed", keep32, with32)
        if operand_cache is not None and key in operand_cache:
            prepacked = operand_cache[key]
        else:
            prepacked = [GalleryOperands.pack(ch, keep32, with32) for ch in chunks]
            if operand_cache is not None:
                operand_cache[key] = prepacked
    ld = (gallery.count + 3) // 4 * 4  # 16-byte aligned rows for the vectorised rank kernel
    grouped = _zeros((probes.count, ld), torch.float32, dev)
    # screen + refine: the tensor-core pass max-reduces its fp16-grade values into `approx`, the refinement its exact
    # float32 values into `grouped`
    approx = _zeros((probes.count, ld), torch.float32, dev) if prec == nat.PREC_FP16_REFINE else None
    if prec == nat.PREC_FP16_REFINE:
        # all variants of a probe in one launch: the refinement then only evaluates the positions that can beat the
        # best variant (a second launch cannot see what a later one will find)
        col_block = max(col_block, 32768)

    def flush(blk: _Block, key: tuple[int, int]) -> None:
        if prepacked is not None:
            _score_block(blk, key, prepacked, offsets, grouped, prec, approx)
            return
        for ch, off in zip(chunks, offsets):
            _score_block(blk, key, [GalleryOperands.pack(ch, keep32, with32)], [off], grouped, prec, approx)

    pending: dict[tuple[int, int], _Block] = {}
    plan = variant_plan(rotations, scales)
    for grp in probes.groups:
        grp.wait()  # a chunk that is still being uploaded: the kernels below queue behind its copy
        n_grp = int(grp.maps.shape[0])
        for s0 in range(0, n_grp, col_block):  # a group wider than a column block is cut, so no block exceeds the cap
            part = grp.maps[s0 : s0 + col_block]
            for rot, scale in plan:
                # a rotation alone is applied by the screening mode's template pack (gather); everything else is
                # materialised when its block is scored
                v = _Variant(part, rot, scale, gather=prec == nat.PREC_FP16_REFINE)
                key = v.shape_hw
                blk = pending.setdefault(key, _Block())
                if blk.ncols and blk.ncols + v.n > col_block:
                    flush(blk, key)
                    blk = pending[key] = _Block()
                blk.maps.append(v)
                blk.ids.append(grp.ids[s0 : s0 + col_block])
                blk.ncols += v.n
        # all variants of this group are in: blocks wide enough to run efficiently go now (the next group may still be
        # on its way from the host), narrow ones keep collecting columns
        if len(probes.groups) > 1:
            for key in [k for k, b in pending.items() if b.ncols >= FLUSH_MIN_COLS]:
                flush(pending.pop(key), key)
    # ragged probe sets leave many narrow blocks (a handful of columns per template shape): those share
    # column tiles through shape buckets instead of running one narrow launch each
    small = {k: b for k, b in pending.items() if b.ncols < bucket_below}
    for key, blk in pending.items():
        if key not in small or len(small) < 2 or prec in (nat.PREC_FP32_SIMT, nat.PREC_FP16X1):
            flush(blk, key)
    if len(small) >= 2 and prec not in (nat.PREC_FP32_SIMT, nat.PREC_FP16X1):
        if prepacked is not None:
            _score_buckets(small, prepacked, offsets, grouped, prec, approx=approx)
        else:
            for ch, off in zip(chunks, offsets):
                _score_buckets(small, [GalleryOperands.pack(ch, False, with32)], [off], grouped, prec, approx=approx)

    # un-group the gallery axis back to the caller's order
    order = torch.cat([ch.ids for ch in chunks])
    if torch.equal(order, torch.arange(gallery.count)):
        return grouped[:, : gallery.count]
    out = torch.empty((probes.count, ld), dtype=torch.float32, device=dev)
    d_order = order.to(torch.int32).to(dev, non_blocking=True)
    nat.check(nat.lib.sir_scatter_columns(_ptr(grouped), probes.count, gallery.count, ld, _ptr(d_order), _ptr(out), ld, _stream()),
              "sir_scatter_columns")
    launch_counter.add()
    return out[:, : gallery.count]


def rank_true_matches(scores: torch.Tensor, true_idx, k: int = 0, g0: int = 0, true_score: torch.Tensor | None = None):
    """(count_gt, count_ge, topk_val, topk_idx, true_score) for ``scores [Q, G_local]``.

    ``rank = 1 + count_gt`` is the reference's ``_get_rank`` (similarity.py:378-386) up to the
    order of exact ties.  ``g0`` is the global index of local column 0; ``true_score`` may be
    supplied when the true match lives on another shard."""
    q, g = int(scores.shape[0]), int(scores.shape[1])
    dev = scores.device
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    ld = int(scores.stride(0))
    tidx = torch.as_tensor(true_idx, dtype=torch.int32).to(dev)
    if true_score is None:
        true_score = torch.empty(q, dtype=torch.float32, device=dev)
        nat.check(nat.lib.sir_true_scores(_ptr(scores), q, g, ld, _ptr(tidx), g0, _ptr(true_score), _stream()), "sir_true_scores")
        launch_counter.add()
    count_gt = torch.empty(q, dtype=torch.int32, device=dev)
    count_ge = torch.empty(q, dtype=torch.int32, device=dev)
    tv = torch.empty((q, max(k, 1)), dtype=torch.float32, device=dev)
    ti = torch.empty((q, max(k, 1)), dtype=torch.int32, device=dev)
    nat.check(
        nat.lib.sir_rank_topk(_ptr(scores), q, g, ld, _ptr(true_score), g0, k, _ptr(count_gt), _ptr(count_ge), _ptr(tv), _ptr(ti), _stream()),
        "sir_rank_topk",
    )
    launch_counter.add()
    return count_gt, count_ge, tv[:, :k], ti[:, :k], true_score


#: host->device bytes of the (probe, gallery) lists of the most recent ``compare`` call
last_h2d_bytes: tuple[int, int] = (0, 0)


def compare(probe_maps, gallery_maps, matching_pairs, rotations=None, scales=None, precision: str = DEFAULT_PRECISION, k: int = 0):
    """Host lists in, (ranks int32 [Q] on host, scores [Q,G] on device, top-k lists) out.

    A ``network.FeatureMapList`` handed over as is (the very object ``get_multiple_feature_maps`` returned) is
    read from its device copies, and its packed gallery operands are kept on the list object
    (``operand_cache``), so a gallery that is compared again -- the next size cluster of ``run.py:17-28`` when
    its features came out of the feature cache -- is neither uploaded nor packed a second time."""
    global last_h2d_bytes

    def ingest(maps, chunk):  # a FeatureMapList keeps its device copies only if it is passed through as is
        if isinstance(maps, MapSet):
            return maps
        return MapSet.from_host(maps if hasattr(maps, "device_copies") else list(maps), chunk=chunk, lazy=True)

    # the gallery goes first (it is needed by the first launch); probes travel in chunks that are scored as they arrive
    gallery, probes = ingest(gallery_maps, None), ingest(probe_maps, PROBE_CHUNK)
    last_h2d_bytes = (probes.h2d_bytes, gallery.h2d_bytes)
    g_total = gallery.count
    tidx = np.asarray(matching_pairs, dtype=np.int64)
    if tidx.size and (tidx.min() < 0 or tidx.max() >= g_total):  # similarity.py:386 raises IndexError there
        raise IndexError("matching_pairs holds an index outside the gallery")
    cache = getattr(gallery_maps, "operand_cache", None) if gallery.h2d_bytes == 0 else None
    scores = score_matrix(probes, gallery, rotations, scales, precision, operand_cache=cache)
    count_gt, _, tv, ti, _ = rank_true_matches(scores, matching_pairs, k)
    ranks = (count_gt + 1).to("cpu").numpy().astype(np.int32)
    return ranks, scores, (tv, ti)
