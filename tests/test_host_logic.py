"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host logic of the
drop-in modules matches the reference-generated goldens / the oracle."""

import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from src.shoeprint_image_retrieval import _native as nat

    header = (ROOT / "include" / "sir.h").read_text()
    declared = set(re.findall(r"\b(sir_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    for name in declared:
        assert hasattr(nat.lib, name)
    assert nat.lib.sir_abi_version() == 3


def test_kpad_matches_layout_rule():
    from src.shoeprint_image_retrieval import _native as nat

    for hm, wm in [(46, 15), (55, 17), (1, 1), (7, 8), (7, 9), (60, 124)]:
        chunks = (wm + 7) // 8
        want = (hm * chunks * 8 + 31) // 32 * 32
        assert nat.lib.sir_template_kpad(hm, wm) == want
    assert nat.lib.sir_template_kpad(0, 5) == 0


def test_argument_errors_are_reported_without_a_gpu():
    import ctypes as C

    from src.shoeprint_image_retrieval import _native as nat

    rc = nat.lib.sir_gallery_pack(None, 1, 1, 8, 8, None, None, None, None, None)
    assert rc == -1
    assert b"null pointer" in nat.lib.sir_last_error()
    with pytest.raises(nat.SirError):
        nat.check(nat.lib.sir_rank_topk(C.c_void_p(8), 1, 1, 1, C.c_void_p(8), 0, 4096, C.c_void_p(8), C.c_void_p(8), None, None, None))


def test_variant_plan_matches_oracle():
    from oracle import variants as ov
    from src.shoeprint_image_retrieval import engine

    cases = [(None, None), ([-15, 9, 180], None), (None, [1.02, 1.08]), ([-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08])]
    for rot, scl in cases:
        assert engine.variant_plan(rot, scl) == [(None if r is None else float(r), None if s is None else float(s)) for r, s in ov.variant_plan(rot, scl)]
    assert len(engine.variant_plan([-15, -9, -3, 3, 9, 15, 180], [1.02, 1.04, 1.08])) == 25
    for h, w, s in [(59, 21, 1.02), (59, 21, 1.08), (50, 19, 1.04), (13, 9, 0.5)]:
        assert engine.scaled_size(h, w, s) == ov.scaled_size(h, w, s)


def test_s_scores_match_reference_vectors(golden, capsys):
    from src.shoeprint_image_retrieval import parse_results as pr

    ranks = [int(r) for r in golden["s_ranks"]]
    tp, tm = int(golden["s_total_prints"]), int(golden["s_total_marks"])
    got = [pr.cmp(ranks, p, tp, tm) for p in (1, 5, 10, 15, 20)]
    np.testing.assert_array_equal(np.array(got), golden["s_values"])
    pr.cmp_all(ranks, tp, tm)
    line = capsys.readouterr().out.strip()
    assert re.fullmatch(r"S1:\d+\.\d\d S5:\d+\.\d\d S10:\d+\.\d\d S15:\d+\.\d\d S20:\d+\.\d\d", line)
    assert line.startswith(f"S1:{got[0] * 100:.2f} ")


def test_load_config_normalises_empty_strings(tmp_path):
    from src.shoeprint_image_retrieval.config import load_config

    cfg = load_config(ROOT / "run.toml")
    assert cfg["comparison"]["rotations"] == [-15, -9, -3, 3, 9, 15, 180]
    assert cfg["model"]["type"] == "EfficientNetV2_M"
    text = (ROOT / "run.toml").read_text()
    text = re.sub(r"rotations\s*=.*", 'rotations = ""', text)
    text = re.sub(r"scales\s*=.*", 'scales = ""', text)
    p = tmp_path / "run.toml"
    p.write_text(text)
    cfg = load_config(p)
    assert cfg["comparison"]["rotations"] is None and cfg["comparison"]["scales"] is None


def test_matching_path_fails_loudly_without_cuda():
    import torch

    from src.shoeprint_image_retrieval import engine

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.MapSet.from_host([np.zeros((2, 8, 8), np.float32)])


def test_feature_map_list_is_a_plain_list_without_device_copies():
    """network.FeatureMapList (what get_multiple_feature_maps returns) behaves as the reference's list of arrays: no device
    copies unless the feature stage attached them, pickles / copies as a plain list."""
    import copy
    import pickle

    import numpy as np

    from src.shoeprint_image_retrieval.network import FeatureMapList

    maps = FeatureMapList([np.zeros((2, 5, 6), np.float32), np.ones((2, 7, 6), np.float32)])
    assert isinstance(maps, list) and len(maps) == 2 and maps.device_copies() is None
    for clone in (pickle.loads(pickle.dumps(maps)), copy.deepcopy(maps)):
        assert type(clone) is list and len(clone) == 2 and np.array_equal(clone[1], maps[1])


@pytest.mark.parametrize("case", ["one", "two"])
def test_dataloader_matches_the_reference_loader(tmp_path, case):
    """What the loader hands to the hot path -- clusters, per-cluster scale and backbone block (incl. the reference's
    _image_extremes quirks), matching pairs and the cropped / LANCZOS-resized uint8 images -- against
    vectors recorded from the reference's own Dataloader on the same generated directories
    (tests/golden/make_golden.py loader_cases; dataloader.py:29-469)."""
    import hashlib
    import sys

    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import synth_dataset

    from src.shoeprint_image_retrieval.dataloader import Dataloader

    with np.load(ROOT / "tests" / "golden" / "reference_loader.npz") as z:
        want = {k: z[k] for k in z.files}
    name, seed, gsizes, qsizes, n_clusters = next(c for c in synth_dataset.CASES if c[0] == case)
    root = tmp_path / name
    synth_dataset.write_dataset(root, seed, gsizes, qsizes)
    loader = Dataloader(synth_dataset.config_for(root, n_clusters))
    assert loader.num_clusters == int(want[f"ld_{name}_nclusters"])
    got = {}
    for k, (marks, prints, pairs, block) in enumerate(loader):
        digest = hashlib.sha256()
        for im in list(marks) + list(prints):
            digest.update(repr(im.shape).encode() + np.ascontiguousarray(im).tobytes())
        got[tuple(sorted(loader.clusters[k]))] = (float(loader.scales[k]), int(block), list(pairs), marks[0], prints[0], digest.hexdigest(),
                                                 [len(marks), len(prints)])
    # the reference's KMeans is unseeded: clusters are matched by their file lists, not by their order
    for k in range(loader.num_clusters):
        files = tuple(str(f) for f in want[f"ld_{name}_c{k}_files"])
        assert files in got, "cluster membership differs from the reference"
        scale, block, pairs, mark0, print0, sha, counts = got[files]
        assert scale == float(want[f"ld_{name}_scales"][k]) and block == int(want[f"ld_{name}_c{k}_block"])
        assert pairs == [int(v) for v in want[f"ld_{name}_c{k}_pairs"]]
        assert counts == [int(v) for v in want[f"ld_{name}_c{k}_counts"]]
        np.testing.assert_array_equal(mark0, want[f"ld_{name}_c{k}_mark0"])
        np.testing.assert_array_equal(print0, want[f"ld_{name}_c{k}_print0"])
        assert sha == str(want[f"ld_{name}_c{k}_sha256"]), "loaded images differ from the reference loader's"


def test_bucket_merge_planner_keeps_every_shape_and_never_costs_more():
    """engine._merge_buckets (host planning of the ragged path): every template shape ends up in exactly one launch, only
    buckets of one mode and orientation are merged, the merged layout covers its members, the modelled cost does not go up,
    and the library's own planner (sir_ncc_cost, host only) accepts every merged bucket shape."""
    from types import SimpleNamespace as NS

    from src.shoeprint_image_retrieval import _native as nat, engine

    rng = np.random.default_rng(5)
    g, hp, wp = 1175, 55, 17
    prec = nat.PREC_FP16_REFINE
    memo: dict = {}

    def cost_of(key):
        if key not in memo:
            mode, flip, bh, bw = key
            memo[key] = engine.plan_cost(mode, g, *((wp, hp) if flip else (hp, wp)), bh, bw)
        return memo[key]

    shapes = {(int(rng.integers(24, 60)), int(rng.integers(10, 22))) for _ in range(160)}
    buckets: dict = {}
    for h, w in sorted(shapes):
        hm, wm = h - 4, w - 4
        flip = (h * 7 + w) % 3 == 0
        oh, ow = (wm, hm) if flip else (hm, wm)
        buckets.setdefault((prec, flip, -(-oh // 8) * 8, -(-ow // 8) * 8), []).append(((h, w), NS(ncols=int(rng.integers(1, 30)))))

    def total(bk):
        out = 0.0
        for key, members in bk.items():
            cols = sum(engine._pad_cols(b.ncols) for _, b in members)
            out += cost_of(key) * engine._tiles_equiv(cols) + engine.BUCKET_FIXED_COST
        return out

    merged = engine._merge_buckets(buckets, cost_of, 8192, 68)
    again = engine._merge_buckets(buckets, cost_of, 8192, 68)
    assert {k: [hw for hw, _ in v] for k, v in merged.items()} == {k: [hw for hw, _ in v] for k, v in again.items()}
    assert sorted(hw for v in merged.values() for hw, _ in v) == sorted(hw for v in buckets.values() for hw, _ in v)
    assert len(merged) < len(buckets)
    for (mode, flip, bh, bw), members in merged.items():
        assert cost_of((mode, flip, bh, bw)) < float("inf")
        for (h, w), _ in members:
            oh, ow = (w - 4, h - 4) if flip else (h - 4, w - 4)
            assert oh <= bh and ow <= bw
            assert any((h, w) in [hw for hw, _ in v] for k, v in buckets.items() if k[:2] == (mode, flip))
    # the bucket-level bound of the model (every tile priced at the bucket's own height) must not exceed the unmerged total
    # by more than the rows-per-tile refinement can give back; the per-tile model itself is monotone by construction
    assert total(merged) <= 1.35 * total(buckets)
    assert engine._tiles_equiv(256) == 1.0 and engine._tiles_equiv(8) == engine.PARTIAL_TILE_FLOOR and engine._tiles_equiv(300) > 1.0
    assert engine._pad_cols(1) == nat.lib.sir_ncc_norm_chunk() and engine._pad_cols(16) == 16


def test_variants_are_described_without_being_generated(monkeypatch):
    """engine._Variant: shape and column count are known at construction (similarity.py:267-274: a rotation keeps the shape,
    a scale gives (int(h*s), int(w*s))), nothing is launched until ``maps`` is read -- the ragged path relies on that to
    queue variant kernels bucket by bucket -- and a scale that shrinks a map to nothing fails immediately."""
    import torch

    from src.shoeprint_image_retrieval import engine

    made = []
    monkeypatch.setattr(engine, "make_variant", lambda maps, rot, scale: made.append((rot, scale)) or torch.zeros(1))
    src = torch.zeros((3, 4, 20, 9))
    plain = engine._Variant(src, None, None, gather=True)
    rotated = engine._Variant(src, 7.0, None, gather=True)
    scaled = engine._Variant(src, 7.0, 1.08, gather=True)
    single_pass = engine._Variant(src, 7.0, None, gather=False)
    assert [v.n for v in (plain, rotated, scaled, single_pass)] == [3, 3, 3, 3]
    assert plain.shape_hw == rotated.shape_hw == single_pass.shape_hw == (20, 9)
    assert scaled.shape_hw == engine.scaled_size(20, 9, 1.08) == (21, 9)
    assert made == []
    assert plain.maps is src and rotated.maps is src and rotated.rot == 7.0 and plain.rot is None
    assert made == []  # a rotation alone is left to the template pack's gather
    assert scaled.rot is None and scaled.maps is not src and made == [(7.0, 1.08)]
    scaled.maps  # noqa: B018 - cached
    assert made == [(7.0, 1.08)]
    assert single_pass.rot is None and single_pass.maps is not src and made[-1] == (7.0, None)
    rotated.materialised()
    assert made[-1] == (7.0, None) and len(made) == 3
    with pytest.raises(ValueError):
        engine._Variant(torch.zeros((1, 4, 20, 9)), None, 0.01)
