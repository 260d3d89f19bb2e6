"""Feature extraction on B200: drop-in for the reference ``network.py``.

Public surface follows the reference (``network.py:90-269``): ``Model(config, block)`` with
``.get_feature_maps(img)`` and ``.get_multiple_feature_maps(images, *, progress=True)``, attributes
``.config .clahe .device .model .transform .transform_rgb``, plus the helpers ``get_output_size``
and ``printmodel``.

What runs where:

* CLAHE (``network.py:108-111,197-208``): grayscale images are equalised on the GPU by
  ``sir_feat_clahe_to_nhwc`` (bit exact with ``cv2.createCLAHE(...).apply``, fused with the
  normalisation); RGB images take the reference's LAB round trip in OpenCV on the host.
  ``SIR_HOST_CLAHE=1`` forces the OpenCV path for grayscale too.
* ToTensor / grayscale repeat / Normalize (``network.py:51-87``) and the truncated backbone
  ``features[:block]`` (``network.py:185-186,234-235``) run through ``libsir.so``: torchvision only
  *defines* the architecture and holds the weights (exactly its role in the reference); at build
  time the module tree is compiled into a flat list of operators with BatchNorm folded, and the
  forward pass is a sequence of ``sir_feat_*`` calls (tcgen05 GEMM convolutions with fused
  bias/SiLU/residual epilogues, depthwise + squeeze-excitation kernels).  No torch operator runs on
  activations.
* Images of equal size are batched (the reference is batch 1, ``network.py:228``); results are
  returned per image as ``[C, h, w]`` float32 numpy arrays like ``network.py:238-244``.

Pretrained weights need torchvision's cached checkpoints (the reference downloads them).  Offline,
pass ``random_init_seed=<int>`` (or set ``SIR_RANDOM_INIT_SEED``) to build the same architecture
with seeded random weights -- what the benchmarks and tests do.

Supported ``config["model"]["type"]``: all 13 strings of the reference -- the EfficientNet family
(B1-B7, V2 S/M/L), VGG16 / VGG19 / VGG19_BN and DenseNet_201 (dense blocks write their growth
channels straight into a shared NHWC buffer, so the concatenation costs nothing).  Unknown strings
raise ``LookupError("Model string not found")`` like ``network.py:181-182``.
"""

from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch
from torch import nn
from tqdm import tqdm

from . import _native as nat
from .engine import _ptr, _require_cuda, _stream, launch_counter

ACT_NONE, ACT_SILU, ACT_RELU = 0, 1, 2

_IMAGENET = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
_VGG16 = ((0.48235, 0.45882, 0.40784), (0.00392156862745098,) * 3)
_HALF = ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))

# model string -> (torchvision constructor, pretrained weights tag, (mean, std))   [network.py:121-177]
_MODELS = {
    "VGG19": ("vgg19", "IMAGENET1K_V1", _IMAGENET),
    "VGG16": ("vgg16", "IMAGENET1K_FEATURES", _VGG16),
    "VGG19_BN": ("vgg19_bn", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNet_B1": ("efficientnet_b1", "IMAGENET1K_V2", _IMAGENET),
    "EfficientNet_B2": ("efficientnet_b2", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNet_B3": ("efficientnet_b3", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNet_B4": ("efficientnet_b4", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNet_B5": ("efficientnet_b5", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNet_B7": ("efficientnet_b7", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNetV2_S": ("efficientnet_v2_s", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNetV2_M": ("efficientnet_v2_m", "IMAGENET1K_V1", _IMAGENET),
    "EfficientNetV2_L": ("efficientnet_v2_l", "IMAGENET1K_V1", _HALF),
    "DenseNet_201": ("densenet201", "IMAGENET1K_V1", _IMAGENET),
}


# --------------------------------------------------------------------------- operator IR

@dataclass
class _Op:
    kind: str            # conv | dwconv | se | affine | maxpool | avgpool | alloc
    src: int             # input tensor id
    dst: int             # output tensor id
    p: dict = field(default_factory=dict)


def _act_code(m: nn.Module | None) -> int:
    if m is None or isinstance(m, nn.Identity):
        return ACT_NONE
    if isinstance(m, nn.SiLU):
        return ACT_SILU
    if isinstance(m, nn.ReLU):
        return ACT_RELU
    raise NotImplementedError(f"activation {type(m).__name__} has no sm_100a kernel yet")


def _fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d | None) -> tuple[torch.Tensor, torch.Tensor]:
    """Conv weight/bias with an eval-mode BatchNorm folded in (float64 arithmetic, float32 result)."""
    w = conv.weight.detach().double()
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    if bn is not None:
        g = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * g[:, None, None, None]
        b = (b - bn.running_mean.detach().double()) * g + bn.bias.detach().double()
    return w.float(), b.float()


class _Compiler:
    """Walks torchvision modules and emits the flat operator list."""

    def __init__(self) -> None:
        self.ops: list[_Op] = []
        self.n_tensors = 1  # tensor 0 = normalised input image
        self.cur = 0

    def _new(self) -> int:
        self.n_tensors += 1
        return self.n_tensors - 1

    def conv(self, conv: nn.Conv2d, bn, act: int, residual: int | None = None, chan_scale: int | None = None,
             into: tuple[int, int] | None = None) -> None:
        """``into = (buffer tensor id, channel offset)``: write the result into a channel slice of an
        existing wider NHWC buffer instead of a fresh tensor (DenseNet concatenation)."""
        if conv.dilation != (1, 1) or conv.padding_mode != "zeros" or conv.padding[0] != conv.padding[1] or conv.stride[0] != conv.stride[1]:
            raise NotImplementedError(f"unsupported convolution {conv}")
        w, b = _fold_bn(conv, bn)
        dst = into[0] if into is not None else self._new()
        common = dict(k=conv.kernel_size[0], kw=conv.kernel_size[1], stride=conv.stride[0], pad=conv.padding[0], act=act, bias=b)
        if conv.groups == 1:
            self.ops.append(_Op("conv", self.cur, dst, dict(common, w=w, cin=conv.in_channels, cout=conv.out_channels,
                                                            residual=residual, chan_scale=chan_scale,
                                                            c_off=None if into is None else into[1])))
        elif conv.groups == conv.in_channels == conv.out_channels and conv.kernel_size[0] == conv.kernel_size[1]:
            self.ops.append(_Op("dwconv", self.cur, dst, dict(common, w=w, c=conv.in_channels)))
        else:
            raise NotImplementedError(f"grouped convolution {conv}")
        self.cur = dst

    def conv_norm_act(self, seq: nn.Sequential, residual: int | None = None, chan_scale: int | None = None) -> None:
        mods = list(seq.children())
        conv = mods[0]
        bn = next((m for m in mods[1:] if isinstance(m, nn.BatchNorm2d)), None)
        actm = next((m for m in mods[1:] if not isinstance(m, nn.BatchNorm2d)), None)
        self.conv(conv, bn, _act_code(actm), residual, chan_scale)

    def affine(self, bn: nn.BatchNorm2d | None, act: int, src_channels: int | None = None) -> None:
        """Standalone BatchNorm (eval) and/or activation; ``src_channels`` = read only the first n channels
        of the (wider) source buffer."""
        scale = shift = None
        if bn is not None:
            g = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            scale, shift = g.float(), (bn.bias.detach().double() - bn.running_mean.detach().double() * g).float()
        dst = self._new()
        self.ops.append(_Op("affine", self.cur, dst, dict(scale=scale, shift=shift, act=act, src_channels=src_channels)))
        self.cur = dst

    def dense_block(self, block: nn.Module) -> None:
        """torchvision ``_DenseBlock``: every layer reads the concatenation of all earlier features
        (BN-ReLU-1x1 conv-BN-ReLU-3x3 conv) and appends ``growth_rate`` channels."""
        layers = list(block.children())
        c0 = layers[0].norm1.num_features
        growth = layers[0].conv2.out_channels
        c_total = c0 + growth * len(layers)
        buf = self._new()
        self.ops.append(_Op("alloc", self.cur, buf, dict(channels=c_total, copy=c0)))
        c_in = c0
        for layer in layers:
            self.cur = buf
            self.affine(layer.norm1, _act_code(layer.relu1), src_channels=c_in)
            self.conv(layer.conv1, layer.norm2, _act_code(layer.relu2))
            self.conv(layer.conv2, None, ACT_NONE, into=(buf, c_in))
            c_in += growth
        self.cur = buf

    def se(self, se: nn.Module) -> int:
        sid = self._new()
        if not isinstance(se.activation, nn.SiLU) or not isinstance(se.scale_activation, nn.Sigmoid):
            raise NotImplementedError("squeeze-excitation with non SiLU/sigmoid activations")
        self.ops.append(_Op("se", self.cur, sid, dict(
            w1=se.fc1.weight.detach().float().flatten(1), b1=se.fc1.bias.detach().float(),
            w2=se.fc2.weight.detach().float().flatten(1), b2=se.fc2.bias.detach().float())))
        return sid

    def module(self, m: nn.Module) -> None:  # noqa: C901, PLR0912
        from torchvision.models.efficientnet import FusedMBConv, MBConv
        from torchvision.ops.misc import Conv2dNormActivation, SqueezeExcitation

        if isinstance(m, Conv2dNormActivation):
            self.conv_norm_act(m)
        elif isinstance(m, (FusedMBConv, MBConv)):
            block_in = self.cur
            res = block_in if m.use_res_connect else None
            layers = list(m.block.children())
            scale_id = None
            for i, layer in enumerate(layers):
                last = i == len(layers) - 1
                if isinstance(layer, SqueezeExcitation):
                    scale_id = self.se(layer)
                else:
                    self.conv_norm_act(layer, residual=res if last else None, chan_scale=scale_id if last else None)
        elif type(m).__name__ == "_DenseBlock":
            self.dense_block(m)
        elif type(m).__name__ == "_Transition":
            self.affine(m.norm, _act_code(m.relu))
            self.conv(m.conv, None, ACT_NONE)
            self.module(m.pool)
        elif isinstance(m, nn.AvgPool2d):
            k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
            st = m.stride if isinstance(m.stride, int) else m.stride[0]
            pd = m.padding if isinstance(m.padding, int) else m.padding[0]
            if pd != 0 or m.ceil_mode:
                raise NotImplementedError(f"unsupported pooling {m}")
            dst = self._new()
            self.ops.append(_Op("avgpool", self.cur, dst, dict(k=k, stride=st)))
            self.cur = dst
        elif isinstance(m, nn.Sequential):
            for child in m.children():
                self.module(child)
        elif isinstance(m, nn.Conv2d):
            self.conv(m, None, ACT_NONE)
        elif isinstance(m, nn.BatchNorm2d):
            self.affine(m, ACT_NONE)
        elif isinstance(m, (nn.ReLU, nn.SiLU)):
            self.affine(None, _act_code(m))
        elif isinstance(m, nn.MaxPool2d):
            k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
            s = m.stride if isinstance(m.stride, int) else m.stride[0]
            pd = m.padding if isinstance(m.padding, int) else m.padding[0]
            if m.ceil_mode or m.dilation not in (1, (1, 1)):
                raise NotImplementedError(f"unsupported pooling {m}")
            dst = self._new()
            self.ops.append(_Op("maxpool", self.cur, dst, dict(k=k, stride=s, pad=pd)))
            self.cur = dst
        elif isinstance(m, (nn.Identity, nn.Dropout)):
            pass
        else:
            raise NotImplementedError(f"module {type(m).__name__} has no sm_100a kernel yet")

    def peephole(self) -> None:
        """Fuse standalone BatchNorm / activation children into the convolution that feeds them
        (VGG-style ``Conv2d, BatchNorm2d, ReLU`` sequences)."""
        out: list[_Op] = []
        for op in self.ops:
            prev = out[-1] if out else None
            if (op.kind == "affine" and prev is not None and prev.kind == "conv" and prev.dst == op.src
                    and prev.p["act"] == ACT_NONE and prev.p["residual"] is None and prev.p["c_off"] is None
                    and op.p["src_channels"] is None):
                if op.p["scale"] is not None:
                    prev.p["w"] = prev.p["w"] * op.p["scale"][:, None, None, None]
                    prev.p["bias"] = prev.p["bias"] * op.p["scale"] + op.p["shift"]
                prev.p["act"] = op.p["act"]
                prev.dst = op.dst
                continue
            if (op.kind == "affine" and prev is not None and prev.kind == "affine" and prev.dst == op.src
                    and prev.p["act"] == ACT_NONE and op.p["scale"] is None and op.p["src_channels"] is None):
                prev.p["act"] = op.p["act"]  # BatchNorm followed by its activation: one pass
                prev.dst = op.dst
                continue
            out.append(op)
        self.ops = out


def _split_fp16(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, int]:
    """x * 2^e (peak in [2^9, 2^10)) as fp16 hi + lo; returns (hi, lo, e)."""
    amax = float(x.abs().max())
    e = 0 if amax == 0 else 10 - int(np.frexp(amax)[1])
    xs = torch.ldexp(x.float(), torch.tensor(e))
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi.contiguous(), lo.contiguous(), e


_COPY_THREADS = 4  # host threads that copy finished feature maps out of the pinned result buffer

# Caches across ``Model`` objects.  ``run.py:17-24`` builds a new ``Model`` and recomputes the features of ALL
# shoeprints for every size cluster: the compiled backbone is kept per (model string, block,
# weights), and the feature maps of an image list per (that key, CLAHE settings, content hash of the images), so a
# cluster that shares its block and scale with an earlier one gets the same ``FeatureMapList`` back -- device copies
# and packed gallery operands included -- without a single kernel launch.  ``SIR_FEATURE_CACHE`` = number of image
# lists kept (default 4, 0 disables).
_PROGRAM_CACHE: dict[tuple, tuple] = {}
_FEATURE_CACHE: "dict[tuple, FeatureMapList]" = {}
feature_cache_stats = {"hits": 0, "misses": 0}


def clear_caches() -> None:
    _PROGRAM_CACHE.clear()
    _FEATURE_CACHE.clear()
    feature_cache_stats.update(hits=0, misses=0)


def _images_digest(images: list[np.ndarray]) -> bytes:
    import hashlib

    h = hashlib.blake2b(digest_size=16)
    for im in images:
        a = np.ascontiguousarray(im)
        h.update(repr((a.shape, a.dtype.str)).encode())
        h.update(a.data)
    return h.digest()


_LAB_TABLES: dict = {}


def _lab_tables(dev: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """OpenCV's 8-bit RGB -> LAB and LAB -> RGB conversions as device tables over all 2^24 pixels (``sir_feat_clahe_rgb_to_nhwc``).
    Both are pure per-pixel functions, so tabulating them with ``cv2.cvtColor`` itself (0.7 s, once per process and device;
    2 x 64 MB) makes the GPU path bit exact with ``network.py:200,204`` whatever OpenCV's fixed-point recipe is."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _LAB_TABLES:
        import cv2

        idx = np.arange(1 << 24, dtype=np.uint32)
        px = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
        out = []
        for code in (cv2.COLOR_RGB2LAB, cv2.COLOR_LAB2RGB):
            conv = cv2.cvtColor(px, code).reshape(-1, 3).astype(np.uint32)
            packed = conv[:, 0] | (conv[:, 1] << 8) | (conv[:, 2] << 16)
            out.append(torch.from_numpy(packed.view(np.int32)).to(dev))
        _LAB_TABLES[key] = tuple(out)
    return _LAB_TABLES[key]


class FeatureMapList(list):
    """What ``get_multiple_feature_maps`` returns: the reference's list of ``[C,h,w]`` float32 arrays
    (``network.py:246-269``) that also remembers the device-resident copies the maps were read back from.
    ``compare_maps`` uses those copies when the list still holds the very same arrays, so the maps do not
    travel host -> device again.  Replacing an element falls back to the host arrays; writing
    into an array in place is not detected."""

    device_groups: list | None = None
    _ids: tuple = ()

    def __init__(self, *args) -> None:
        super().__init__(*args)
        #: packed gallery operands of these maps (filled by ``engine.compare`` when the list is used as a gallery)
        self.operand_cache: dict = {}

    def attach_device_copies(self, chunks: list) -> None:
        by_shape: dict[tuple, tuple[list[int], list[torch.Tensor]]] = {}
        for shp, idx, maps in chunks:
            ids, parts = by_shape.setdefault(shp, ([], []))
            ids.extend(idx)
            parts.append(maps)
        groups = []
        for ids, parts in by_shape.values():
            if len(parts) == 1:
                groups.append((parts[0], ids))
                continue
            whole = torch.empty((sum(int(t.shape[0]) for t in parts), *parts[0].shape[1:]), dtype=parts[0].dtype, device=parts[0].device)
            at = 0
            for t in parts:  # device-to-device copies (cudaMemcpyAsync), no torch kernel
                whole[at : at + int(t.shape[0])].copy_(t)
                at += int(t.shape[0])
            groups.append((whole, ids))
        self.device_groups = groups
        self._ids = tuple(id(a) for a in self)

    def __reduce__(self):  # pickles (and deep-copies) as the plain list of arrays; the device copies stay behind
        return (list, (list(self),))

    def device_copies(self) -> list | None:
        """``[(tensor [n,C,h,w], list indices)]`` if every element is still the array that was returned, else None."""
        if self.device_groups is None or len(self) != len(self._ids) or any(id(a) != i for a, i in zip(self, self._ids)):
            return None
        return self.device_groups


class _Program:
    """Compiled backbone: device-resident weights + the executor."""

    def __init__(self, modules: list[nn.Module], device: torch.device) -> None:
        comp = _Compiler()
        for m in modules:
            comp.module(m)
        comp.peephole()
        self.ops = comp.ops
        self.n_tensors = comp.n_tensors
        self.launch_log: list[str] | None = None  # profiling aid: one label per kernel launched by run()
        self.out_id = comp.cur if not self.ops else self.ops[-1].dst
        self.device = device
        for op in self.ops:
            p = op.p
            if op.kind == "conv":
                cout, cin, k, kw = p["cout"], p["cin"], p["k"], p["kw"]
                bn = int(nat.lib.sir_feat_conv_tile_n(cout))
                rows = (cout + bn - 1) // bn * bn
                w = p["w"].permute(0, 2, 3, 1)  # [cout][ky][kx][cin]
                # |out| <= amax_in * max_n sum|w[n]| + max|bias| (SiLU/ReLU do not grow it): a-priori scale of emitted planes
                p["bound_mult"] = float(p["w"].double().abs().flatten(1).sum(1).max()) * (1.0 + 1e-5)
                p["bound_add"] = float(p["bias"].abs().max()) * (1.0 + 1e-5)
                implicit = cin % 8 == 0  # any stride: strided patches come through TMA element strides
                p["direct"] = cin == 3 and k == 3 and kw == 3 and cout % 8 == 0 and cout <= 256 and p["c_off"] is None
                if p["direct"]:  # the stem: float32 on the CUDA cores, weights [27][cout]
                    p["w_direct"] = w.reshape(cout, 27).t().contiguous().to(device)
                if implicit:  # K = (tap, channel) with the channels of a tap padded to the K step
                    bk = 32 if cin % 32 == 0 else 16 if cin % 16 == 0 else 32
                    cp = (cin + bk - 1) // bk * bk
                    wm = torch.zeros((rows, k * kw, cp), dtype=torch.float32)
                    wm[:cout, :, :cin] = w.reshape(cout, k * kw, cin)
                    wm = wm.reshape(rows, k * kw * cp)
                else:  # explicit im2col matrix, K = (tap, channel) packed and padded to 32
                    bk = 32
                    kdim = k * kw * cin
                    wm = torch.zeros((rows, (kdim + 31) // 32 * 32), dtype=torch.float32)
                    wm[:cout, :kdim] = w.reshape(cout, kdim)
                hi, lo, e = _split_fp16(wm)
                bias = torch.zeros((cout + 3) // 4 * 4, dtype=torch.float32)
                bias[:cout] = p["bias"]
                p.update(whi=hi.to(device), wlo=lo.to(device), w_exp=e, kp=int(wm.shape[1]), rows=rows, bias_d=bias.to(device),
                         implicit=implicit, bk=bk)
                del p["w"]
            elif op.kind == "dwconv":
                p["bound_mult"] = float(p["w"].double().abs().flatten(1).sum(1).max()) * (1.0 + 1e-5)
                p["bound_add"] = float(p["bias"].abs().max()) * (1.0 + 1e-5)
                p.update(w_d=p["w"][:, 0].permute(1, 2, 0).contiguous().to(device), bias_d=p["bias"].to(device))
                del p["w"]
            elif op.kind == "se":
                p["w2"] = p["w2"].t()  # fc2 as [S][C]: coalesced over channels
                for key in ("w1", "b1", "w2", "b2"):
                    p[key] = p[key].contiguous().to(device)
            elif op.kind == "affine" and p["scale"] is not None:
                p["scale"], p["shift"] = p["scale"].to(device), p["shift"].to(device)

        # Which convolution outputs are handed on as fp16 operand planes (written by the producer's epilogue, no
        # split pass) and which are (also) needed in float32.
        def takes_planes(u: _Op) -> bool:
            return u.kind == "conv" and u.p["implicit"] and u.p["chan_scale"] is None

        for op in self.ops:
            if op.kind != "conv":
                continue
            t = op.dst
            plane_users = [u for u in self.ops if u.src == t and takes_planes(u)]
            other = t == self.out_id or any(
                (u.src == t and not takes_planes(u)) or u.p.get("residual") == t or u.p.get("chan_scale") == t for u in self.ops)
            ok = op.p["c_off"] is None and op.p["cout"] % 8 == 0 and os.environ.get("SIR_NO_PLANE_CHAIN", "") != "1"
            op.p["emit_planes"] = bool(plane_users) and ok
            op.p["emit_f32"] = other or not op.p["emit_planes"]

        # MBConv: depthwise 3x3 -> SqueezeExcitation -> 1x1 projection.  The depthwise kernel writes the projection's operand
        # planes itself and the SE scale is folded into per-image projection weights, so nothing re-reads the activation.
        for op in self.ops:
            if op.kind != "dwconv":
                continue
            t = op.dst
            users = [u for u in self.ops if u.src == t or u.p.get("residual") == t or u.p.get("chan_scale") == t]
            convs = [u for u in users if u.kind == "conv"]
            fast = op.p["k"] == 3 and op.p["kw"] == 3 and op.p["pad"] == 1 and op.p["stride"] in (1, 2) and op.p["c"] % 8 == 0
            ok = (len(convs) == 1 and all(u.kind in ("conv", "se") and u.src == t for u in users) and t != self.out_id
                  and convs[0].p["implicit"] and convs[0].p["k"] == 1 and convs[0].p["kw"] == 1 and convs[0].p["pad"] == 0
                  and convs[0].p["chan_scale"] is not None and os.environ.get("SIR_NO_PLANE_CHAIN", "") != "1")
            op.p["emit_planes"] = bool(ok and fast)
            if op.p["emit_planes"]:
                convs[0].p["scaled_planes"] = True

    @staticmethod
    def _out_hw(h: int, w: int, k: int, kw: int, s: int, pd: int) -> tuple[int, int]:
        return (h + 2 * pd - k) // s + 1, (w + 2 * pd - kw) // s + 1

    def _packed_weights(self, p: dict, geom: tuple, st: C.c_void_p) -> tuple[torch.Tensor, int, int]:
        """The layer's weights as per-stage shared-memory images for the tile the library plans for this shape
        (packed on the device the first time a tile shape is needed, then cached)."""
        tile_n, granule, nbytes = C.c_int(), C.c_int(), C.c_longlong()
        nat.check(nat.lib.sir_feat_conv_plan(*geom, p["bk"], p["cout"], 0, C.byref(tile_n), C.byref(granule), C.byref(nbytes)), "sir_feat_conv_plan")
        key = (tile_n.value, granule.value)
        packs = p.setdefault("packs", {})
        if key not in packs:
            buf = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            nat.check(nat.lib.sir_feat_conv_pack_weights(_ptr(p["whi"]), _ptr(p["wlo"]), p["rows"], p["kp"], key[0], key[1], _ptr(buf), st),
                      "sir_feat_conv_pack_weights")
            packs[key] = buf
        return packs[key], key[0], key[1]

    def _scaled_weights(self, p: dict, geom: tuple, scale: torch.Tensor, st: C.c_void_p) -> tuple[torch.Tensor, int, int]:
        """One packed weight set per image with the squeeze-excitation scale ``[B,C]`` folded in."""
        tile_n, granule, nbytes = C.c_int(), C.c_int(), C.c_longlong()
        nat.check(nat.lib.sir_feat_conv_plan(*geom, p["bk"], p["cout"], 1, C.byref(tile_n), C.byref(granule), C.byref(nbytes)), "sir_feat_conv_plan")
        b, cin = geom[0], geom[3]
        buf = torch.empty((b, nbytes.value), dtype=torch.uint8, device=self.device)
        nat.check(nat.lib.sir_feat_conv_scale_weights(_ptr(p["whi"]), _ptr(p["wlo"]), p["rows"], p["kp"], p["kp"] // (p["k"] * p["kw"]), cin,
                                                      _ptr(scale), b, tile_n.value, granule.value, _ptr(buf), st), "sir_feat_conv_scale_weights")
        return buf, tile_n.value, granule.value

    def run(self, x0: torch.Tensor, amax0: torch.Tensor) -> torch.Tensor:  # noqa: C901, PLR0915
        """x0: normalised input [B,H,W,3] float32 NHWC; amax0: 1-element tensor with its max |x|."""
        dev = self.device
        st = _stream()
        tensors: dict[int, torch.Tensor] = {0: x0}
        amax = torch.zeros(self.n_tensors, dtype=torch.float32, device=dev)
        amax[0:1] = amax0
        last_use = {}
        for i, op in enumerate(self.ops):
            last_use[op.src] = i
            if op.p.get("c_off") is not None:
                last_use[op.dst] = i  # the dense block's buffer stays alive while layers write into it
            for key in ("residual", "chan_scale"):
                if op.p.get(key) is not None:
                    last_use[op.p[key]] = i

        log = self.launch_log
        squeezed = {op.src for op in self.ops if op.kind == "se"}
        pooled: dict[int, tuple[torch.Tensor, int]] = {}
        planes: dict[int, tuple[torch.Tensor, torch.Tensor]] = {}  # tensor id -> fp16 hi/lo operand planes from its producer
        exps = torch.zeros(self.n_tensors, dtype=torch.int32, device=dev)

        def aptr(tid: int) -> C.c_void_p:
            return C.c_void_p(amax.data_ptr() + 4 * tid)

        def describe(op: _Op, n0: int) -> None:
            q = op.p
            text = op.kind
            if op.kind == "conv":
                text = f"conv k{q['k']} s{q['stride']} {q['cin']}->{q['cout']}" + (" +planes" if q["emit_planes"] else "") + ("" if q["emit_f32"] else " -f32")
            elif op.kind == "dwconv":
                text = f"dw k{q['k']} s{q['stride']}"
            log.extend([text] * (launch_counter.n - n0))

        for i, op in enumerate(self.ops):
            p = op.p
            src = tensors[op.src]
            b, h, w, c = (int(v) for v in src.shape)
            n_before = launch_counter.n
            if op.kind == "conv":
                ho, wo = self._out_hw(h, w, p["k"], p["kw"], p["stride"], p["pad"])
                m = b * ho * wo
                cs = tensors[p["chan_scale"]] if p["chan_scale"] is not None else None
                res = tensors[p["residual"]] if p["residual"] is not None else None
                if p["c_off"] is None:
                    out = torch.empty((b, ho, wo, p["cout"]), dtype=torch.float32, device=dev if p["emit_f32"] else "meta")
                    out_ptr, ldc = (_ptr(out) if p["emit_f32"] else None), p["cout"]
                else:  # growth channels of a dense layer go straight into the block's buffer
                    out = tensors[op.dst]
                    out_ptr, ldc = C.c_void_p(out.data_ptr() + 4 * p["c_off"]), int(out.shape[3])
                ohi = olo = exp_out = None
                if p["emit_planes"]:
                    ohi = torch.empty((b, ho, wo, p["cout"]), dtype=torch.float16, device=dev)
                    olo = torch.empty_like(ohi)
                    planes[op.dst] = (ohi, olo)
                    exp_out = C.c_void_p(exps.data_ptr() + 4 * op.dst)
                if p["direct"] and cs is None and res is None:
                    nat.check(nat.lib.sir_feat_conv_c3k3(_ptr(src), aptr(op.src), b, h, w, p["stride"], p["pad"], _ptr(p["w_direct"]),
                                                         _ptr(p["bias_d"]), p["cout"], p["act"], out_ptr, aptr(op.dst), _ptr(ohi), _ptr(olo),
                                                         exp_out, p["bound_mult"], p["bound_add"], st), "sir_feat_conv_c3k3")
                    launch_counter.add()
                    tensors[op.dst] = out
                    for tid in [t for t, lu in last_use.items() if lu == i and t != self.out_id]:
                        tensors.pop(tid, None)
                        planes.pop(tid, None)
                    if log is not None:
                        describe(op, n_before)
                    continue
                exp_in = None
                per_image = 0
                if p.get("scaled_planes") and op.src in planes:
                    # depthwise planes + SE scale folded into one weight set per image (no pass over the activation)
                    ahi, alo = planes[op.src]
                    exp_in = C.c_void_p(exps.data_ptr() + 4 * op.src)
                    geom = (b, h, w, c, p["k"], p["kw"], p["pad"], p["stride"])
                    per_image = 1
                elif p["implicit"] and cs is None and op.src in planes:  # operand planes came from the producer's epilogue
                    ahi, alo = planes[op.src]
                    exp_in = C.c_void_p(exps.data_ptr() + 4 * op.src)
                    geom = (b, h, w, c, p["k"], p["kw"], p["pad"], p["stride"])
                    launch_counter.add(-1)
                elif p["implicit"]:  # split once into fp16 hi/lo NHWC planes; the kernel gathers the taps itself
                    ahi = torch.empty((b, h, w, c), dtype=torch.float16, device=dev)
                    alo = torch.empty_like(ahi)
                    nat.check(nat.lib.sir_feat_im2col_split(_ptr(src), aptr(op.src), b, h, w, c, 1, 1, 1, 0, _ptr(cs), c, _ptr(ahi), _ptr(alo), st),
                              "sir_feat_im2col_split")
                    geom = (b, h, w, c, p["k"], p["kw"], p["pad"], p["stride"])
                else:
                    ahi = torch.empty((m, p["kp"]), dtype=torch.float16, device=dev)
                    alo = torch.empty_like(ahi)
                    nat.check(nat.lib.sir_feat_im2col_split(_ptr(src), aptr(op.src), b, h, w, c, p["k"], p["kw"], p["stride"], p["pad"],
                                                            _ptr(cs), p["kp"], _ptr(ahi), _ptr(alo), st), "sir_feat_im2col_split")
                    geom = (1, 1, m, p["kp"], 1, 1, 0, 1)
                if per_image:
                    wpack, tile_n, granule = self._scaled_weights(p, geom, cs, st)
                else:
                    wpack, tile_n, granule = self._packed_weights(p, geom, st)
                nat.check(nat.lib.sir_feat_conv(_ptr(ahi), _ptr(alo), aptr(op.src), *geom, p["bk"], _ptr(wpack), tile_n, granule, per_image,
                                                p["cout"], p["w_exp"], _ptr(p["bias_d"]), _ptr(res), p["act"],
                                                out_ptr, ldc, aptr(op.dst), exp_in, _ptr(ohi), _ptr(olo), exp_out, p["bound_mult"],
                                                p["bound_add"], aptr(p["residual"]) if p["residual"] is not None else None, st),
                          "sir_feat_conv")
                launch_counter.add(2)
            elif op.kind == "dwconv":
                ho, wo = self._out_hw(h, w, p["k"], p["kw"], p["stride"], p["pad"])
                emit = p.get("emit_planes", False)
                out = torch.empty((b, ho, wo, c), dtype=torch.float32, device="meta" if emit else dev)
                part = None
                if op.dst in squeezed:  # a SqueezeExcitation follows: its global sum is produced here
                    parts = int(nat.lib.sir_feat_dwconv_pool_parts(p["k"], p["stride"], c, ho, wo))
                    part = torch.empty((b, parts, c), dtype=torch.float32, device=dev)
                    pooled[op.dst] = (part, parts)
                ohi = olo = exp_out = None
                if emit:  # the projection reads these planes; the SE scale goes into its weights
                    ohi = torch.empty((b, ho, wo, c), dtype=torch.float16, device=dev)
                    olo = torch.empty_like(ohi)
                    planes[op.dst] = (ohi, olo)
                    exp_out = C.c_void_p(exps.data_ptr() + 4 * op.dst)
                nat.check(nat.lib.sir_feat_dwconv(_ptr(src), b, h, w, c, p["k"], p["stride"], p["pad"], _ptr(p["w_d"]),
                                                  _ptr(p["bias_d"]), p["act"], None if emit else _ptr(out), aptr(op.dst), _ptr(part),
                                                  aptr(op.src), _ptr(ohi), _ptr(olo), exp_out, p["bound_mult"], p["bound_add"], st),
                          "sir_feat_dwconv")
                launch_counter.add()
            elif op.kind == "se":
                if op.src in pooled:
                    part, parts = pooled.pop(op.src)
                else:
                    part, parts = torch.empty((b, 1, c), dtype=torch.float32, device=dev), 1
                    nat.check(nat.lib.sir_feat_pool_sum(_ptr(src), b, h * w, c, _ptr(part), st), "sir_feat_pool_sum")
                    launch_counter.add()
                avg = torch.empty((b, c), dtype=torch.float32, device=dev)
                out = torch.empty((b, c), dtype=torch.float32, device=dev)
                nat.check(nat.lib.sir_feat_se_scale(_ptr(part), b, parts, h * w, c, int(p["w1"].shape[0]), _ptr(p["w1"]), _ptr(p["b1"]),
                                                    _ptr(p["w2"]), _ptr(p["b2"]), _ptr(avg), _ptr(out), st), "sir_feat_se_scale")
                launch_counter.add(2)
            elif op.kind == "affine":
                cs_ = p["src_channels"] or c
                out = torch.empty((b, h, w, cs_), dtype=torch.float32, device=dev)
                nat.check(nat.lib.sir_feat_affine_act(_ptr(src), b * h * w, cs_, c, cs_, _ptr(p["scale"]), _ptr(p["shift"]), p["act"],
                                                      _ptr(out), aptr(op.dst), st), "sir_feat_affine_act")
                launch_counter.add()
            elif op.kind == "alloc":
                out = torch.empty((b, h, w, p["channels"]), dtype=torch.float32, device=dev)
                nat.check(nat.lib.sir_feat_affine_act(_ptr(src), b * h * w, p["copy"], c, p["channels"], None, None, ACT_NONE,
                                                      _ptr(out), aptr(op.dst), st), "sir_feat_affine_act")
                launch_counter.add()
            elif op.kind == "avgpool":
                ho, wo = (h - p["k"]) // p["stride"] + 1, (w - p["k"]) // p["stride"] + 1
                out = torch.empty((b, ho, wo, c), dtype=torch.float32, device=dev)
                nat.check(nat.lib.sir_feat_avgpool2d(_ptr(src), b, h, w, c, p["k"], p["stride"], _ptr(out), aptr(op.dst), st),
                          "sir_feat_avgpool2d")
                launch_counter.add()
            elif op.kind == "maxpool":
                ho, wo = self._out_hw(h, w, p["k"], p["k"], p["stride"], p["pad"])
                out = torch.empty((b, ho, wo, c), dtype=torch.float32, device=dev)
                nat.check(nat.lib.sir_feat_maxpool(_ptr(src), b, h, w, c, p["k"], p["stride"], p["pad"], _ptr(out), aptr(op.dst), st),
                          "sir_feat_maxpool")
                launch_counter.add()
            else:  # pragma: no cover
                raise AssertionError(op.kind)
            tensors[op.dst] = out
            if log is not None:
                describe(op, n_before)
            for tid in [t for t, lu in last_use.items() if lu == i and t != self.out_id]:
                tensors.pop(tid, None)
                planes.pop(tid, None)
        return tensors[self.out_id]


# --------------------------------------------------------------------------- public API

def printmodel(model: nn.Module, input_shape: tuple[int, int, int, int] = (1, 3, 1968, 5872)) -> None:
    """Print the architecture with torchinfo (``network.py:16-29``); debug helper."""
    from torchinfo import summary

    with torch.no_grad():
        print(summary(model, input_shape))


def get_output_size(model: "Model", input_shape: tuple[int, int, int, int]) -> torch.Size:
    """Shape ``[1, C, h, w]`` of the feature maps for an input of ``input_shape`` (``network.py:32-48``),
    computed from the compiled operator list without running the network."""
    _, _, h, w = input_shape
    c: int | None = 3
    block_channels = 0
    for op in model.program.ops:
        p = op.p
        if op.kind == "conv":
            h, w = _Program._out_hw(h, w, p["k"], p["kw"], p["stride"], p["pad"])
            c = p["cout"]
        elif op.kind == "dwconv":
            h, w = _Program._out_hw(h, w, p["k"], p["kw"], p["stride"], p["pad"])
        elif op.kind == "maxpool":
            h, w = _Program._out_hw(h, w, p["k"], p["k"], p["stride"], p["pad"])
        elif op.kind == "avgpool":
            h, w = (h - p["k"]) // p["stride"] + 1, (w - p["k"]) // p["stride"] + 1
        if op.kind == "conv" and p["c_off"] is not None:
            c = None  # channel count of a dense block comes from its buffer
        if op.kind == "alloc":
            block_channels = p["channels"]
        if c is None:
            c = block_channels
    return torch.Size((input_shape[0], c, h, w))


class Model:
    """Truncated pretrained backbone + pre-processing (reference ``network.py:90-269``)."""

    def __init__(self, config: dict, block: int, *, random_init_seed: int | None = None) -> None:
        import cv2
        from torchvision import models

        self.config = config
        self.clahe = cv2.createCLAHE(
            clipLimit=config["model"]["clahe_clip_limit"],
            tileGridSize=tuple(config["model"]["clahe_tile_grid_size"]),
        )
        self.device = _require_cuda()
        model_str = config["model"]["type"]
        if model_str not in _MODELS:
            raise LookupError("Model string not found")  # network.py:181-182
        ctor, tag, (mean, std) = _MODELS[model_str]
        if random_init_seed is None and os.environ.get("SIR_RANDOM_INIT_SEED"):
            random_init_seed = int(os.environ["SIR_RANDOM_INIT_SEED"])
        weights_key = tag if random_init_seed is None else f"seed{random_init_seed}"
        self._key = (model_str, int(block), weights_key, self.device.index)
        if self._key not in _PROGRAM_CACHE:
            if random_init_seed is None:
                net = getattr(models, ctor)(weights=tag)  # downloads / reads torchvision's cache like the reference
            else:
                gen_state = torch.random.get_rng_state()
                torch.manual_seed(random_init_seed)
                net = getattr(models, ctor)(weights=None)
                torch.random.set_rng_state(gen_state)
            net.eval()
            layers = list(net.features.children())[:block]  # network.py:185
            if len(_PROGRAM_CACHE) >= 4:
                _PROGRAM_CACHE.pop(next(iter(_PROGRAM_CACHE)))
            _PROGRAM_CACHE[self._key] = (nn.Sequential(*layers).eval(), _Program(layers, self.device))
        # the module list is kept for introspection (weights live there, on the host)
        self.model, self.program = _PROGRAM_CACHE[self._key]
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.transform = self._host_transform(gray=True)
        self.transform_rgb = self._host_transform(gray=False)
        self.max_batch_bytes = 4 << 30  # cap on the live tensors of one layer
        self.max_batch = 64
        self._host_clahe = os.environ.get("SIR_HOST_CLAHE", "") == "1"
        self._copy_stream: torch.cuda.Stream | None = None
        self._copy_pool = ThreadPoolExecutor(max_workers=_COPY_THREADS)
        self._pinned_bufs: dict[tuple, torch.Tensor] = {}

    def _host_transform(self, *, gray: bool):
        mean, std = np.array(self.mean, np.float32), np.array(self.std, np.float32)

        def apply(img: np.ndarray) -> torch.Tensor:
            x = torch.from_numpy(np.ascontiguousarray(img)).float().div(255)
            x = x[None].repeat(3, 1, 1) if gray else x.permute(2, 0, 1)
            return (x - torch.from_numpy(mean)[:, None, None]) / torch.from_numpy(std)[:, None, None]

        return apply

    def _clahe(self, img: np.ndarray) -> np.ndarray:
        """CLAHE on the luminance (``network.py:197-208``): RGB goes through LAB."""
        import cv2

        if img.ndim == 3:
            lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
            l_ch, a_ch, b_ch = cv2.split(lab)
            return cv2.cvtColor(cv2.merge((self.clahe.apply(l_ch), a_ch, b_ch)), cv2.COLOR_LAB2RGB)
        return self.clahe.apply(img)

    # ---- device path -----------------------------------------------------------------------
    def _forward_uint8(self, batch: np.ndarray, *, apply_clahe: bool = False) -> torch.Tensor:
        """uint8 images ``[B,H,W]`` or ``[B,H,W,3]`` -> feature maps ``[B,C,h,w]`` on the device.
        ``apply_clahe``: the (grayscale) batch is raw and CLAHE runs on the GPU; otherwise it is already equalised."""
        d_img = torch.from_numpy(np.ascontiguousarray(batch)).to(self.device, non_blocking=True)
        return self._forward_device(d_img, apply_clahe=apply_clahe)

    def _forward_device(self, d_img: torch.Tensor, *, apply_clahe: bool) -> torch.Tensor:
        """The same for a uint8 batch already on the device."""
        b, h, w = (int(v) for v in d_img.shape[:3])
        in_ch = 1 if d_img.ndim == 3 else 3
        x0 = torch.empty((b, h, w, 3), dtype=torch.float32, device=self.device)
        amax0 = torch.zeros(1, dtype=torch.float32, device=self.device)
        mean = (C.c_float * 3)(*self.mean)
        std = (C.c_float * 3)(*self.std)
        if apply_clahe:
            tx, ty = (int(v) for v in self.config["model"]["clahe_tile_grid_size"])
            clip = float(self.config["model"]["clahe_clip_limit"])
            lut = torch.empty((b, tx * ty, 256), dtype=torch.uint8, device=self.device)
            if in_ch == 1:
                nat.check(nat.lib.sir_feat_clahe_to_nhwc(_ptr(d_img), b, h, w, clip, tx, ty, mean, std, _ptr(lut), None, _ptr(x0), _ptr(amax0),
                                                         _stream()), "sir_feat_clahe_to_nhwc")
                launch_counter.add(2)
            else:  # RGB prints: LAB round trip through the tabulated OpenCV conversions (network.py:199-204)
                rgb2lab, lab2rgb = _lab_tables(self.device)
                l_plane = torch.empty((b, h, w), dtype=torch.uint8, device=self.device)
                ab_plane = torch.empty((b, h, w), dtype=torch.int16, device=self.device)
                nat.check(nat.lib.sir_feat_clahe_rgb_to_nhwc(_ptr(d_img), b, h, w, clip, tx, ty, mean, std, _ptr(rgb2lab), _ptr(lab2rgb),
                                                             _ptr(l_plane), _ptr(ab_plane), _ptr(lut), None, _ptr(x0), _ptr(amax0), _stream()),
                          "sir_feat_clahe_rgb_to_nhwc")
                launch_counter.add(3)
        else:
            nat.check(nat.lib.sir_feat_image_to_nhwc(_ptr(d_img), b, h, w, in_ch, mean, std, _ptr(x0), _ptr(amax0), _stream()),
                      "sir_feat_image_to_nhwc")
            launch_counter.add()
        y = self.program.run(x0, amax0)
        bo, ho, wo, co = (int(v) for v in y.shape)
        out = torch.empty((bo, co, ho, wo), dtype=torch.float32, device=self.device)
        nat.check(nat.lib.sir_feat_nhwc_to_nchw(_ptr(y), bo, ho * wo, co, _ptr(out), _stream()), "sir_feat_nhwc_to_nchw")
        launch_counter.add()
        return out

    def _pinned(self, kind: str, slot: int, shape: tuple, dtype: torch.dtype) -> torch.Tensor:
        """Reusable page-locked staging buffer (two slots per kind and shape).  Both slots are created together: page-locking
        tens of megabytes takes milliseconds and would otherwise land inside the second chunk of every first call."""
        shape = tuple(int(v) for v in shape)
        key = (kind, slot, shape, dtype)
        if key not in self._pinned_bufs:
            for k in [k for k in self._pinned_bufs if k[0] == kind]:
                del self._pinned_bufs[k]
            for sl in (0, 1):
                self._pinned_bufs[(kind, sl, shape, dtype)] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._pinned_bufs[key]

    def _batch_limit(self, h: int, w: int) -> int:
        """Images per forward pass: the largest (input + fp16 operand planes + output) of any layer within
        ``max_batch_bytes``, at most ``max_batch`` (late layers need ~64 images to fill 148 SMs)."""
        worst = 1
        hh, ww, c = h, w, 3
        for op in self.program.ops:
            p = op.p
            if op.kind in ("conv", "dwconv"):
                ho, wo = _Program._out_hw(hh, ww, p["k"], p["kw"], p["stride"], p["pad"])
                if op.kind == "conv":
                    operand = hh * ww * c * 4 if p["implicit"] else ho * wo * p["kp"] * 4
                    worst = max(worst, hh * ww * c * 4 + operand + ho * wo * p["cout"] * 4)
                    c = p["cout"]
                else:
                    worst = max(worst, (hh * ww + ho * wo) * c * 4)
                hh, ww = ho, wo
            elif op.kind == "maxpool":
                hh, ww = _Program._out_hw(hh, ww, p["k"], p["k"], p["stride"], p["pad"])
            elif op.kind == "avgpool":
                hh, ww = (hh - p["k"]) // p["stride"] + 1, (ww - p["k"]) // p["stride"] + 1
            elif op.kind == "alloc":
                c = p["channels"]
                worst = max(worst, hh * ww * c * 8)
        return max(1, min(self.max_batch, int(self.max_batch_bytes // worst)))

    def get_feature_maps(self, img: np.ndarray) -> np.ndarray:
        """One image (uint8 ``[H,W]`` or ``[H,W,3]``) -> ``[C,h,w]`` float32 (``network.py:210-244``)."""
        if img.ndim in (2, 3) and img.dtype == np.uint8 and not self._host_clahe:
            out = self._forward_uint8(img[None], apply_clahe=True)
        else:
            out = self._forward_uint8(self._clahe(img)[None])
        return out.cpu().numpy().squeeze()  # squeeze like network.py:244

    def get_multiple_feature_maps(self, images: list[np.ndarray], *, progress: bool = True) -> list[np.ndarray]:
        """List of images -> list of feature maps, same order (``network.py:246-269``).  Images of
        equal shape are pushed through the backbone as one batch."""
        slots = int(os.environ.get("SIR_FEATURE_CACHE", "4"))
        cache_key = None
        if slots > 0 and len(images) > 0:
            model_cfg = self.config["model"]
            cache_key = (self._key, float(model_cfg["clahe_clip_limit"]), tuple(model_cfg["clahe_tile_grid_size"]), self._host_clahe,
                         len(images), _images_digest(images))
            hit = _FEATURE_CACHE.get(cache_key)
            if hit is not None and hit.device_copies() is not None:
                feature_cache_stats["hits"] += 1
                _FEATURE_CACHE[cache_key] = _FEATURE_CACHE.pop(cache_key)  # most recently used last
                if progress:
                    with tqdm(total=len(images)) as bar:
                        bar.update(len(images))
                return hit
            feature_cache_stats["misses"] += 1
        results: list[Any] = [None] * len(images)
        by_shape: dict[tuple, list[int]] = {}
        for i, im in enumerate(images):
            by_shape.setdefault(tuple(im.shape), []).append(i)
        bar = tqdm(total=len(images)) if progress else None
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        pending: tuple | None = None  # (image indices, pinned host maps, copy-done event, device maps)
        device_chunks: list[tuple[tuple, list[int], torch.Tensor]] = []

        def collect(pend: tuple) -> None:
            """Copy a finished chunk out of its pinned buffer: one pageable block per chunk, filled by a few threads
            (the cost is first-touch page faults, which numpy's copy takes with the GIL released)."""
            chunk, host, done, _dev = pend
            done.synchronize()
            arr = host.numpy()
            block = np.empty(arr.shape, arr.dtype)
            n = len(chunk)
            parts = min(_COPY_THREADS, n)

            def fill(a: int, b: int) -> None:
                block[a:b] = arr[a:b]

            jobs = [self._copy_pool.submit(fill, n * t // parts, n * (t + 1) // parts) for t in range(parts)]
            for j in jobs:
                j.result()
            for j, i in enumerate(chunk):
                results[i] = block[j].squeeze()
            if bar is not None:
                bar.update(n)

        # Two-deep pipeline: while the GPU works on chunk i the host stages chunk i+1 into pinned memory and copies the
        # maps of chunk i-1 out of the pinned result buffer; results leave the device on a side stream.
        step = 0
        for shp, idx in by_shape.items():
            limit = self._batch_limit(shp[0], shp[1])
            gpu_clahe = not self._host_clahe and all(images[i].dtype == np.uint8 for i in idx)
            for s in range(0, len(idx), limit):
                chunk = idx[s : s + limit]
                stage = self._pinned("in", step % 2, (limit, *shp), torch.uint8)[: len(chunk)]
                if gpu_clahe:
                    np.stack([images[i] for i in chunk], out=stage.numpy())
                else:
                    np.stack([self._clahe(images[i]) for i in chunk], out=stage.numpy())
                d_img = torch.empty(stage.shape, dtype=torch.uint8, device=self.device)
                d_img.copy_(stage, non_blocking=True)
                maps = self._forward_device(d_img, apply_clahe=gpu_clahe)
                ready = torch.cuda.Event()
                ready.record(main)
                host = self._pinned("out", step % 2, (limit, *maps.shape[1:]), torch.float32)[: len(chunk)]
                done = torch.cuda.Event()
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ready)
                    host.copy_(maps, non_blocking=True)
                    done.record(self._copy_stream)
                maps.record_stream(self._copy_stream)
                if pending is not None:
                    collect(pending)
                pending = (chunk, host, done, maps)
                device_chunks.append((tuple(maps.shape[1:]), chunk, maps))
                step += 1
        if pending is not None:
            collect(pending)
        if bar is not None:
            bar.close()
        out = FeatureMapList(results)
        budget = float(os.environ.get("SIR_DEVICE_MAP_CACHE_GB", "16")) * 2**30
        if sum(m.numel() * 4 for _, _, m in device_chunks) <= budget:
            out.attach_device_copies(device_chunks)
            if cache_key is not None:
                while len(_FEATURE_CACHE) >= slots:
                    _FEATURE_CACHE.pop(next(iter(_FEATURE_CACHE)))
                _FEATURE_CACHE[cache_key] = out
        return out
