"""Dataset loading for ``run.py``: behaviour-compatible with the reference ``dataloader.py``.

Off the hot path (file I/O and image decode) but needed for
the unchanged entry point.  What is kept exactly, because it decides the inputs of the hot path:

* directory layout ``<dir>/Gallery`` + ``<dir>/Query``; ids parsed from file names per dataset type
  (``dataloader.py:245-250``); FID-300 matches come from ``label_table.csv`` (``:99-107``);
* crop by ``floor(size * ratio)`` per edge, then LANCZOS resize to ``int(size * scale)``
  (``:218-237``); gallery images are re-loaded per cluster at that cluster's scale (``:86-90``);
* per-cluster scale and backbone block from the recursive rule of ``_find_best_scale``
  (``:366-419``), fed by ``_image_extremes`` INCLUDING its two quirks (``height, width =
  image.size`` although PIL returns (width, height), and the ``elif`` that lets an image update only
  one extreme, ``:446,462``) so that the selected block matches the reference;
* clusters whose scales differ by at most ``cluster_minimise_tolerance`` and share a block merge
  (``:329-364``).

Optional (``SIR_DEVICE_RESIZE=1``, off by default): the LANCZOS resize of a cluster's images runs on the GPU
(``engine.resize_images_lanczos`` -> ``sir_image_resize_lanczos``, bit exact with Pillow's 8-bit resampler); decode and crop stay
on the host threads.  It is off by default because the loader hands numpy arrays to the caller, so the resized images
travel back to the host and the round trip costs more than Pillow's own resize saves.

Deliberate differences: KMeans is seeded (the reference's is not, ``:284``, so its clustering
varies run to run); files are decoded by a thread pool and every file is loaded (the reference's
process chunking drops or misplaces files when ``len % n_processes`` is 1, ``:137-146``).
"""

from __future__ import annotations

import csv
import os
from concurrent.futures import ThreadPoolExecutor
from math import floor
from pathlib import Path
import numpy as np
from PIL import Image


def _file_id(name: str, dataset_type: str) -> int:
    if dataset_type == "Impress":
        return int(name.split("_")[0].split(".")[0])
    if dataset_type == "WVU2019":
        return int(name[:3])
    if dataset_type == "FID-300":
        return int(name[:-4])
    raise ValueError(f"unknown dataset type {dataset_type!r}")


class Dataloader:
    """Iterator over size clusters: ``(shoemark images, shoeprint images, matching ids, block)``."""

    def __init__(self, config: dict) -> None:
        self.config = config
        self.dataset_dir = Path(config["dataset"]["dir"])
        self.shoeprint_dir = self.dataset_dir / "Gallery"
        self.shoemark_dir = self.dataset_dir / "Query"
        self.shoeprint_files = os.listdir(self.shoeprint_dir)
        self.shoemark_files = os.listdir(self.shoemark_dir)
        print(
            "The dataset contains: \n",
            f"    {len(self.shoeprint_files)} reference shoeprints\n",
            f"    {len(self.shoemark_files)} shoemarks",
        )
        clustered = self._cluster_images_by_size(self.shoemark_dir, config["dataset"]["n_clusters"])
        self.scales, self.blocks, self.clusters = self._minimise_clusters(clustered)
        self.num_clusters = len(self.clusters)
        self._current_cluster = 0

    def __iter__(self) -> "Dataloader":
        return self

    def __next__(self) -> tuple[list[np.ndarray], list[np.ndarray], list[int], int]:
        if self._current_cluster >= self.num_clusters:
            raise StopIteration
        k = self._current_cluster
        marks, mark_ids = self._load_images(self.clusters[k], self.shoemark_dir, self.scales[k])
        prints, print_ids = self._load_images(self.shoeprint_files, self.shoeprint_dir, self.scales[k])
        if self.config["dataset"]["type"] != "FID-300":
            # many shoemarks may share one shoeprint (WVU2019): index of the shoeprint with the same id
            pairs = [print_ids.index(i) for i in mark_ids]
        else:
            with (self.dataset_dir / "label_table.csv").open() as fh:
                table = {int(row[0]): int(row[1]) for row in csv.reader(fh) if row}
            pairs = [table[i] - 1 for i in mark_ids]
        self._current_cluster += 1
        return marks, prints, pairs, self.blocks[k]

    # ------------------------------------------------------------------ loading
    def _load_one(self, directory: Path, name: str, scale: float, resize: bool = True) -> np.ndarray:
        crop = self.config["dataset"]["crop"]
        with Image.open(directory / name) as image:
            ch, cw = floor(image.height * crop[0]), floor(image.width * crop[1])
            image = image.crop((cw, ch, image.width - cw, image.height - ch))
            if not resize:
                return np.array(image)
            size = (int(image.width * scale), int(image.height * scale))
            return np.array(image.resize(size, Image.Resampling.LANCZOS))

    def _load_images(self, image_files: list[str], image_directory: Path, scale: float) -> tuple[list[np.ndarray], list[int]]:
        image_files.sort()  # in place, like the reference (dataloader.py:133): gallery order = name order
        workers = max(1, min(int(self.config["dataset"]["n_processes"]), os.cpu_count() or 1, len(image_files) or 1))
        device_resize = os.environ.get("SIR_DEVICE_RESIZE", "") == "1"
        with ThreadPoolExecutor(workers) as pool:
            images = list(pool.map(lambda n: self._load_one(image_directory, n, scale, not device_resize), image_files))
        if device_resize:  # decode + crop on the host, LANCZOS on the GPU (bit exact with Pillow)
            from . import engine

            usable = [im.dtype == np.uint8 and im.ndim in (2, 3) for im in images]
            sizes = [(int(im.shape[0] * scale), int(im.shape[1] * scale)) for im in images]
            if all(usable):
                images = engine.resize_images_lanczos(images, sizes)
            else:  # 16-bit / palette images: Pillow's own path
                images = [np.array(Image.fromarray(im).resize((hw[1], hw[0]), Image.Resampling.LANCZOS)) for im, hw in zip(images, sizes)]
        ids = [_file_id(n, self.config["dataset"]["type"]) for n in image_files]
        return images, ids

    # ------------------------------------------------------------------ clustering / scale selection
    def _cluster_images_by_size(self, image_dir: Path, n_clusters: int) -> dict[int, list[str]]:
        from sklearn.cluster import KMeans

        names = os.listdir(image_dir)
        sizes = []
        for name in names:
            with Image.open(image_dir / name) as image:
                sizes.append([min(image.size)])
        n_clusters = max(1, min(n_clusters, len({s[0] for s in sizes})))
        labels = KMeans(n_clusters=n_clusters, n_init=10, random_state=0).fit(sizes).labels_
        clusters: dict[int, list[str]] = {}
        for name, label in zip(names, labels):
            clusters.setdefault(int(label), []).append(name)
        return clusters

    def _image_extremes(self, image_files: list[str], image_directory: Path) -> tuple[tuple[str, int], tuple[str, int]]:
        crop = self.config["dataset"]["crop"]
        big_name, big = "", 0
        small_name, small = "", 2**31 - 1
        for name in image_files:
            with Image.open(image_directory / name) as image:
                height, width = image.size  # sic: PIL gives (width, height); kept for block parity (App. D7)
            height -= floor(height * crop[0] * 2)
            width -= floor(width * crop[1] * 2)
            if max(width, height) > big:
                big_name, big = name, max(width, height)
            elif min(width, height) < small:  # sic: elif, an image updates one extreme only
                small_name, small = name, min(width, height)
        return (big_name, big), (small_name, small)

    def _find_best_scale(self, smallest_dim: int, largest_dim: int, minimum_dim: int, block: int) -> tuple[float, int]:
        """Recursive scale / block choice ("Algorithm 1", ``dataloader.py:366-419``)."""
        model = self.config["model"]
        maximum_dim, end_block, skip = model["maximum_dim"], model["end_block"], model["skip_blocks"]
        scale: float = 1
        if smallest_dim < minimum_dim:
            if block > end_block:
                block -= 1
                while block in skip:
                    block -= 1
                return self._find_best_scale(smallest_dim, largest_dim, int(minimum_dim / 2), block)
            return 1, block
        if largest_dim > maximum_dim:
            scale = maximum_dim / largest_dim
            if smallest_dim * scale < minimum_dim:
                if block > end_block:
                    block -= 1
                    while block in skip and block != end_block:
                        block -= 1
                else:
                    scale = minimum_dim / smallest_dim
        return scale, block

    def _minimise_clusters(self, clusters: dict[int, list[str]]) -> tuple[list[float], list[int], list[list[str]]]:
        tol = self.config["dataset"]["cluster_minimise_tolerance"]
        scales: list[float] = []
        blocks: list[int] = []
        groups: list[list[str]] = []
        big_print, small_print = self._image_extremes(self.shoeprint_files, self.shoeprint_dir)
        for files in clusters.values():
            big_mark, small_mark = self._image_extremes(files, self.shoemark_dir)
            smallest = min(small_mark[1], small_print[1])
            largest = max(big_mark[1], big_print[1])
            scale, block = self._find_best_scale(smallest, largest, self.config["model"]["minimum_dim"], self.config["model"]["start_block"])
            # the reference merges into the FIRST earlier cluster within tolerance, and only if its block matches
            hit = next((i for i, s in enumerate(scales) if abs(scale - s) <= tol), None)
            if hit is not None and blocks[hit] == block:
                groups[hit] += files
            else:
                scales.append(scale)
                blocks.append(block)
                groups.append(list(files))
        return scales, blocks, groups
