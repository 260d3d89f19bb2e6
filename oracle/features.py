"""Oracle: feature-map extraction on the CPU with PyTorch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``Model.get_feature_maps`` (``src/shoeprint_image_retrieval/network.py:210-244``):
CLAHE (cv2) -> ToTensor -> grayscale repeat -> Normalize -> ``nn.Sequential(features[:block])`` in
eval mode under ``no_grad`` -> numpy ``[C,h,w]``.  The arithmetic is torch / torchvision's
(third party; the reference pins torch 1.13.1 / torchvision 0.14.1, ``pyproject.toml:11-12``).
"""

from __future__ import annotations

import numpy as np
import torch
from torch import nn


def normalise(img: np.ndarray, mean, std) -> torch.Tensor:
    """ToTensor (/255), repeat a grayscale image to 3 channels, Normalize (network.py:51-87)."""
    x = torch.from_numpy(np.ascontiguousarray(img)).float().div(255)
    x = x[None].repeat(3, 1, 1) if img.ndim == 2 else x.permute(2, 0, 1)
    mean_t = torch.tensor(mean, dtype=torch.float32)[:, None, None]
    std_t = torch.tensor(std, dtype=torch.float32)[:, None, None]
    return (x - mean_t) / std_t


def feature_maps(layers: nn.Sequential, img_after_clahe: np.ndarray, mean, std, dtype=torch.float32) -> np.ndarray:
    """Forward one pre-processed (CLAHE'd) uint8 image through ``layers`` on the CPU."""
    x = normalise(img_after_clahe, mean, std)[None].to(dtype)
    with torch.no_grad():
        y = layers.to(dtype)(x)
    return y[0].float().numpy()


def images_per_second(layers: nn.Sequential, images: list[np.ndarray], mean, std) -> tuple[float, int]:
    """Time the batch-1 loop of ``get_multiple_feature_maps`` (network.py:246-269) on the host cores."""
    import time

    layers = layers.float().eval()
    feature_maps(layers, images[0], mean, std)  # warm-up
    t0 = time.perf_counter()
    for im in images:
        feature_maps(layers, im, mean, std)
    return len(images) / (time.perf_counter() - t0), torch.get_num_threads()
