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


def test_run_py_matches_oracle_pipeline(tmp_path, monkeypatch, capsys):
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
    toml_text = (ge.ROOT / "run.toml").read_text()
    toml_text = re.sub(r'dir = ".*?"', f'dir = "{data}/"', toml_text)
    toml_text = toml_text.replace('type = "Impress"', 'type = "FID-300"')
    toml_text = re.sub(r"n_clusters = \d+", "n_clusters = 2", toml_text)
    cfg_path = tmp_path / "run.toml"
    cfg_path.write_text(toml_text)
    monkeypatch.setenv("SIR_RANDOM_INIT_SEED", "11")

    run.main(str(cfg_path))
    out = capsys.readouterr().out
    lines = [ln for ln in out.splitlines() if ln.startswith("S1:")]
    assert lines, out
    got_ranks = [int(m) for m in re.findall(r"true match ranked (\d+)", out + capsys.readouterr().err)]

    # oracle pipeline on the same inputs
    config = load_config(cfg_path)
    loader = Dataloader(config)
    all_ranks, total_q = [], len(loader.shoemark_files)
    for marks, prints, pairs, block in loader:
        model = network.Model(config, block)  # same seeded weights; used here only as the weight holder + CLAHE
        feats_q = [ofeat.feature_maps(model.model, model._clahe(im), model.mean, model.std) for im in marks]
        feats_g = [ofeat.feature_maps(model.model, model._clahe(im), model.mean, model.std) for im in prints]
        ranks, scores = ocmp.compare_maps_oracle(feats_q, feats_g, pairs, config["comparison"]["rotations"], config["comparison"]["scales"])
        gpu_scores = similarity.last_result["scores"].cpu().numpy() if loader.num_clusters == 1 else None
        if gpu_scores is not None:
            err = np.abs(gpu_scores - scores) / np.maximum(np.abs(scores), 1e-3)
            assert err.max() < 5e-4, f"end-to-end score error {err.max():.2e}"
        for q in range(len(marks)):
            lo, hi = ocmp.rank_interval(scores[q], pairs[q], rel_tol=1e-3)
            all_ranks.append((lo, hi))
    if loader.num_clusters == 1 and got_ranks:
        for r, (lo, hi) in zip(got_ranks, all_ranks):
            assert lo <= r <= hi
    want = ocmp.s_scores([lo for lo, _ in all_ranks], len(loader.shoeprint_files), total_q)
    assert re.fullmatch(r"S1:\d+\.\d\d S5:\d+\.\d\d S10:\d+\.\d\d S15:\d+\.\d\d S20:\d+\.\d\d", lines[-1])
    if loader.num_clusters == 1 and all(lo == hi for lo, hi in all_ranks):
        assert lines[-1] == " ".join(f"{k}:{v:.2f}" for k, v in want.items())
