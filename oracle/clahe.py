"""Oracle: OpenCV CLAHE for 8-bit single-channel images, restated in numpy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference calls ``cv2.createCLAHE(clipLimit, tileGridSize).apply(img)`` (``network.py:108-111,
197-208``; opencv-python is third party, pinned 4.10.0.84 in ``uv.lock``, 4.13 installed here).  Its
published algorithm (modules/imgproc/src/clahe.cpp):

1. if the image size is not a multiple of the tile grid, extend it to the right/bottom with
   BORDER_REFLECT_101 (the LUTs are computed on the extended image, the output keeps the size);
2. per tile: 256-bin histogram, clip at ``max(1, int(clipLimit * tileArea / 256))``, redistribute the
   clipped mass (``clipped // 256`` to every bin, the remainder one count every ``max(256 // residual, 1)``
   bins from bin 0), cumulative sum, ``lut[i] = saturate(round_half_even(sum_i * 255 / tileArea))`` with
   the scale in float32;
3. per pixel: bilinear interpolation (float32, products and sums rounded separately) between the LUTs of
   the four surrounding tile centres, ``round_half_even`` to uint8.

Checked bit-for-bit against cv2 in ``tests/test_oracle_clahe.py``.
"""

from __future__ import annotations

import numpy as np


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    idx = np.where(idx >= n, 2 * n - 2 - idx, idx)
    return idx


def clahe_luts(img: np.ndarray, clip_limit: float, tiles_x: int, tiles_y: int) -> tuple[np.ndarray, int, int]:
    h, w = img.shape
    if w % tiles_x == 0 and h % tiles_y == 0:
        ext = img
    else:
        eh = h + (tiles_y - h % tiles_y)
        ew = w + (tiles_x - w % tiles_x)
        ys = _reflect101(np.arange(eh), h)
        xs = _reflect101(np.arange(ew), w)
        ext = img[np.ix_(ys, xs)]
    th, tw = ext.shape[0] // tiles_y, ext.shape[1] // tiles_x
    area = th * tw
    lut_scale = np.float32(255.0) / np.float32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((tiles_y * tiles_x, 256), dtype=np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            tile = ext[ty * th : (ty + 1) * th, tx * tw : (tx + 1) * tw]
            hist = np.bincount(tile.reshape(-1), minlength=256).astype(np.int64)
            if clip > 0:
                over = np.maximum(hist - clip, 0)
                clipped = int(over.sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            csum = np.cumsum(hist).astype(np.float32)
            val = np.rint(csum * lut_scale)  # cvRound: half to even; float32 product
            luts[ty * tiles_x + tx] = np.clip(val, 0, 255).astype(np.uint8)
    return luts, th, tw


def clahe(img: np.ndarray, clip_limit: float = 2.0, tiles: tuple[int, int] = (8, 8)) -> np.ndarray:
    """``cv2.createCLAHE(clip_limit, tiles).apply(img)`` for a uint8 2-D image."""
    tiles_x, tiles_y = int(tiles[0]), int(tiles[1])
    luts, th, tw = clahe_luts(img, clip_limit, tiles_x, tiles_y)
    h, w = img.shape
    f32 = np.float32
    inv_tw, inv_th = f32(1.0) / f32(tw), f32(1.0) / f32(th)
    txf = np.arange(w, dtype=f32) * inv_tw - f32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(f32)).astype(f32)
    xa1 = (f32(1.0) - xa).astype(f32)
    tx2 = np.minimum(tx1 + 1, tiles_x - 1)
    tx1 = np.maximum(tx1, 0)
    tyf = np.arange(h, dtype=f32) * inv_th - f32(0.5)
    ty1 = np.floor(tyf).astype(np.int64)
    ya = (tyf - ty1.astype(f32)).astype(f32)
    ya1 = (f32(1.0) - ya).astype(f32)
    ty2 = np.minimum(ty1 + 1, tiles_y - 1)
    ty1 = np.maximum(ty1, 0)
    v = img.astype(np.int64)
    l11 = luts[(ty1[:, None] * tiles_x + tx1[None, :]), v].astype(f32)
    l12 = luts[(ty1[:, None] * tiles_x + tx2[None, :]), v].astype(f32)
    l21 = luts[(ty2[:, None] * tiles_x + tx1[None, :]), v].astype(f32)
    l22 = luts[(ty2[:, None] * tiles_x + tx2[None, :]), v].astype(f32)
    top = (l11 * xa1[None, :] + l12 * xa[None, :]).astype(f32)
    bot = (l21 * xa1[None, :] + l22 * xa[None, :]).astype(f32)
    res = (top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)
    return np.clip(np.rint(res.astype(f32)), 0, 255).astype(np.uint8)
