"""End-to-end drop-in check on a generated FID-300-shaped directory: the unchanged entry point
(run.py -> Dataloader -> Model -> compare_maps -> cmp_all) on the GPU against the CPU oracle pipeline
(oracle.features with the same seeded random-init backbone + oracle.compare)."""

import os
import re

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _make_dataset(root, n_gallery=6, n_query=5, seed=0):
    rng = np.random.default_rng(seed)
    (root / "Gallery").mkdir()
    (root / "Query").mkdir()
    prints = []
    for i in range(1, n_gallery + 1):
        base = rng.integers(0, 256, size=(74, 35)).astype(np.float32)
        img = np.kron(base, np.ones((8, 8), np.float32))[:586, :270]
        img = np.clip(img * 0.6 + rng.normal(60, 25, img.shape), 0, 255).astype(np.uint8)
        prints.append(img)
        Image.fromarray(img).save(root / "Gallery" / f"{i:05d}.png")
    rows = []
    for q in range(1, n_query + 1):
        g = int(rng.integers(0, n_gallery))
        h, w = int(rng.integers(300, 586)), int(rng.integers(180, 270))
        y0, x0 = int(rng.integers(0, 586 - h + 1)), int(rng.integers(0, 270 - w + 1))
        crop = prints[g][y0 : y0 + h, x0 : x0 + w].astype(np.float32)
        crop = np.clip(crop + rng.normal(0, 10, crop.shape), 0, 255).astype(np.uint8)
        Image.fromarray(crop).save(root / "Query" / f"{q:05d}.png")
        rows.append(f"{q},{g + 1}")
    (root / "label_table.csv").write_text("\n".join(rows) + "\n")


def _write_config(tmp_path, data, n_clusters=1):
    import __graft_entry__ as ge

    toml_text = (ge.ROOT / "run.toml").read_text()
    toml_text = re.sub(r'dir = ".*?"', f'dir = "{data}/"', toml_text)
    toml_text = toml_text.replace('type = "Impress"', 'type = "FID-300"')
    toml_text = re.sub(r"n_clusters = \d+", f"n_clusters = {n_clusters}", toml_text)
    cfg_path = tmp_path / "run.toml"
    cfg_path.write_text(toml_text)
    return cfg_path


def test_run_py_matches_oracle_pipeline(tmp_path, monkeypatch, capsys):
    """The unchanged entry point on one size cluster (n_clusters = 1, so every assertion below always runs):
    * the feature maps never travel host -> device again between the stages (h2d_bytes == 0 for both lists);
    * compare stage on IDENTICAL maps: GPU scores vs the CPU oracle on the very maps the GPU feature stage produced,
      1e-4 relative, ranks inside their 1e-4 intervals, printed S-scores equal when every interval is a singleton;
    * feature stage: GPU maps vs the torch modules evaluated on the CPU, relative L2 < 1e-4 per map."""
    import __graft_entry__ as ge

    ge.build()
    import run
    from oracle import compare as ocmp
    from oracle import features as ofeat
    from src.shoeprint_image_retrieval import network, similarity
    from src.shoeprint_image_retrieval.config import load_config
    from src.shoeprint_image_retrieval.dataloader import Dataloader

    data = tmp_path / "fid"
    data.mkdir()
    _make_dataset(data)
    cfg_path = _write_config(tmp_path, data, n_clusters=1)
    monkeypatch.setenv("SIR_RANDOM_INIT_SEED", "11")
    network.clear_caches()

    run.main(str(cfg_path))
    captured = capsys.readouterr()
    out = captured.out
    lines = [ln for ln in out.splitlines() if ln.startswith("S1:")]
    assert len(lines) == 1, out
    got_ranks = [int(m) for m in re.findall(r"true match ranked (\d+)", out + captured.err)]
    assert similarity.last_result["h2d_bytes"] == (0, 0), "compare_maps re-uploaded maps the feature stage had left on the device"
    gpu_scores = similarity.last_result["scores"].cpu().numpy()

    config = load_config(cfg_path)
    loader = Dataloader(config)
    assert loader.num_clusters == 1
    marks, prints, pairs, block = next(loader)
    assert len(got_ranks) == len(marks)
    model = network.Model(config, block)
    hits0 = network.feature_cache_stats["hits"]
    feats_q = model.get_multiple_feature_maps(marks, progress=False)
    feats_g = model.get_multiple_feature_maps(prints, progress=False)
    assert network.feature_cache_stats["hits"] == hits0 + 2  # the maps run.main() computed: same arrays, no kernels

    # compare stage on identical maps
    rot, scl = config["comparison"]["rotations"], config["comparison"]["scales"]
    ranks, scores = ocmp.compare_maps_oracle(list(feats_q), list(feats_g), pairs, rot, scl)
    err = np.abs(gpu_scores - scores) / np.maximum(np.abs(scores), 1e-3)
    assert err.max() < 1e-4, f"compare-stage score error {err.max():.2e}"
    intervals = [ocmp.rank_interval(scores[q], pairs[q], rel_tol=1e-4) for q in range(len(marks))]
    for r, (lo, hi) in zip(got_ranks, intervals):
        assert lo <= r <= hi
    assert re.fullmatch(r"S1:\d+\.\d\d S5:\d+\.\d\d S10:\d+\.\d\d S15:\d+\.\d\d S20:\d+\.\d\d", lines[0])
    if all(lo == hi for lo, hi in intervals):
        want = ocmp.s_scores([lo for lo, _ in intervals], len(loader.shoeprint_files), len(loader.shoemark_files))
        assert lines[0] == " ".join(f"{k}:{v:.2f}" for k, v in want.items())

    # feature stage
    for img, got in list(zip(marks, feats_q))[:3] + list(zip(prints, feats_g))[:3]:
        want = ofeat.feature_maps(model.model, model._clahe(img), model.mean, model.std)
        rel = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert got.shape == want.shape and rel < 1e-4, rel


def test_second_cluster_reuses_gallery_features_and_operands(tmp_path, monkeypatch, capsys):
    """run.py recomputes the features of ALL shoeprints for every size cluster (run.py:17-24).  When two clusters share
    block and scale -- the loader hands out the same gallery images again -- the second pass must not run the backbone
    on the gallery, upload or re-pack it: run.main() with a loader that yields two clusters over one gallery."""
    import __graft_entry__ as ge

    ge.build()
    import run
    from src.shoeprint_image_retrieval import engine, network, similarity
    from src.shoeprint_image_retrieval.dataloader import Dataloader

    data = tmp_path / "fid"
    data.mkdir()
    _make_dataset(data, n_gallery=5, n_query=4)
    cfg_path = _write_config(tmp_path, data, n_clusters=1)
    monkeypatch.setenv("SIR_RANDOM_INIT_SEED", "11")
    network.clear_caches()

    class TwoClusters(Dataloader):
        """Both halves of the shoemarks as separate clusters with the same scale and block (what _minimise_clusters
        leaves behind when an earlier cluster of another block sits within the scale tolerance, dataloader.py:329-364)."""

        def __init__(self, config):
            super().__init__(config)
            files = sorted(self.clusters[0])
            self.clusters = [files[: len(files) // 2], files[len(files) // 2 :]]
            self.scales, self.blocks = [self.scales[0]] * 2, [self.blocks[0]] * 2
            self.num_clusters = 2

    seen = []
    real_compare = similarity.compare_maps

    def spy(marks, prints, pairs, config):
        before = engine.launch_counter.n
        ranks = real_compare(marks, prints, pairs, config)
        seen.append({"prints": prints, "h2d": similarity.last_result["h2d_bytes"], "packs": dict(prints.operand_cache),
                     "launches": engine.launch_counter.n - before})
        return ranks

    monkeypatch.setattr(run, "Dataloader", TwoClusters)
    monkeypatch.setattr(run, "compare_maps", spy)
    run.main(str(cfg_path))
    assert len(seen) == 2
    assert seen[0]["prints"] is seen[1]["prints"], "cluster 2 recomputed the gallery feature maps"
    assert network.feature_cache_stats["hits"] >= 1
    assert seen[0]["h2d"] == (0, 0) and seen[1]["h2d"] == (0, 0)
    key = next(iter(seen[0]["packs"]))
    assert seen[1]["packs"][key] is seen[0]["packs"][key], "cluster 2 re-packed the gallery operands"
    assert len([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("S1:")]) == 2
