"""GPU parity of the fused correlation kernel (tcgen05 and SIMT paths) and of the drop-in
``compare_maps`` against the CPU oracle and the reference-generated golden vectors.

Tolerance (BASELINE.json north_star): 1e-4 relative on scores; ranks identical except where the
oracle's own score gaps fall below that tolerance (rank intervals)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

REL_TOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as ge

    ge.build()
    from src.shoeprint_image_retrieval import engine

    return engine


def _check(got, want, tol=REL_TOL):
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
    assert err.max() <= tol, f"max rel err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}: got {got.flat[err.argmax()]} want {want.flat[err.argmax()]}"


CASES = [
    # (C, gallery hw, G, Q, min_frac, rotations, scales)
    (4, (16, 12), 5, 4, 1.0, None, None),            # uniform shapes, one patch
    (3, (24, 13), 4, 5, 0.6, None, None),            # ragged probes, 2x2 patches, Wm from 1 to 9
    (6, (30, 25), 3, 3, 0.8, [-9, 15, 180], None),   # rotations, nkc up to 3
    (5, (21, 15), 4, 3, 0.7, None, [1.04, 1.2]),     # scales (template can outgrow the gallery)
    (4, (19, 14), 3, 3, 0.7, [-15, 9], [1.08]),      # both: 1 + (R+1)*S variants
    (2, (40, 9), 2, 2, 1.0, None, None),             # tall and thin
    (2, (9, 44), 2, 2, 1.0, None, None),             # short and wide (6 chunks per row)
    (2, (70, 40), 2, 2, 1.0, None, None),            # large template: several E segments per channel
    (2, (68, 132), 1, 1, 1.0, None, None),           # 1024x2048-input sized maps (config 5 shape), 16 chunks/row
]


@pytest.mark.parametrize("precision,tol", [("fp32_simt", 2e-5), ("fp16x3", REL_TOL), ("fp16_fp8c", REL_TOL), ("fp16_refine", 2e-5)])
@pytest.mark.parametrize("case", CASES)
def test_scores_match_oracle(eng, case, precision, tol):
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    c, (h, w), g, q, frac, rot, scl = case
    gallery = synth.make_gallery(11, g, c, h, w)
    probes, pairs = synth.make_probes(12, gallery, q, min_frac=frac)
    ranks, scores, _ = eng.compare(probes, gallery, pairs, rot, scl, precision=precision)
    want_ranks, want = ocmp.compare_maps_oracle(probes, gallery, pairs, rot, scl, method="fast")
    _check(scores.cpu().numpy(), want, tol)
    for i in range(q):
        lo, hi = ocmp.rank_interval(want[i], pairs[i])
        assert lo <= ranks[i] <= hi


def test_fp16x1_is_close_but_looser(eng):
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    gallery = synth.make_gallery(21, 6, 8, 22, 16)
    probes, pairs = synth.make_probes(22, gallery, 5, min_frac=0.8)
    _, scores, _ = eng.compare(probes, gallery, pairs, None, None, precision="fp16x1")
    _, want = ocmp.compare_maps_oracle(probes, gallery, pairs, None, None)
    _check(scores.cpu().numpy(), want, 2e-3)


def test_many_columns_and_ragged_gallery(eng):
    """More than one 256-column tile, several gallery shapes, gallery order restored."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    g1 = synth.make_gallery(31, 3, 3, 14, 11)
    g2 = synth.make_gallery(32, 2, 3, 12, 13)
    gallery = [g1[0], g2[0], g1[1], g2[1], g1[2]]
    rng = np.random.default_rng(33)
    probes = [np.ascontiguousarray(g1[i % 3] + 0.3 * rng.standard_normal(g1[0].shape).astype(np.float32)) for i in range(300)]
    pairs = [[0, 2, 4][i % 3] for i in range(300)]
    ranks, scores, _ = eng.compare(probes, gallery, pairs, None, None, precision="fp16x3")
    got = scores.cpu().numpy()
    sub = list(range(0, 300, 37))
    _, want = ocmp.compare_maps_oracle([probes[i] for i in sub], gallery, [pairs[i] for i in sub])
    _check(got[sub], want)
    assert np.all(ranks == 1)


def test_simt_and_tensor_core_agree_at_reference_shapes(eng):
    """FID-300 block-4 shape (80x59x21) and WVU-like block-6 shape (176x50x19), no oracle (too
    slow on CPU at this size): the two independent GPU evaluations must agree to 1e-4."""
    from src.shoeprint_image_retrieval import synth

    for c, h, w in [(80, 59, 21), (176, 50, 19)]:
        gal = synth.device_gallery(41, 9, c, h, w)
        prb, pairs = synth.device_probes(42, gal, 7)
        ps, gs = eng.MapSet.from_device(prb), eng.MapSet.from_device(gal)
        a = eng.score_matrix(ps, gs, [-5, 5], None, "fp32_simt").cpu().numpy()
        b = eng.score_matrix(ps, gs, [-5, 5], None, "fp16x3").cpu().numpy()
        _check(b, a)
        c8 = eng.score_matrix(ps, gs, [-5, 5], None, "fp16_fp8c").cpu().numpy()
        _check(c8, a)
        rf = eng.score_matrix(ps, gs, [-5, 5], None, "fp16_refine").cpu().numpy()
        _check(rf, a, 1e-5)
        print(f"C={c}: max rel err fp16x3 {np.max(np.abs(b - a) / np.maximum(a, 1e-3)):.2e}, fp16_fp8c {np.max(np.abs(c8 - a) / np.maximum(a, 1e-3)):.2e}")
        assert np.all(a.argmax(1) == pairs.cpu().numpy())


@pytest.mark.parametrize("mode", ["none", "rot", "scl", "both"])
def test_compare_maps_dropin_matches_reference_vectors(eng, golden, mode, capsys):
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval.similarity import compare_maps

    gallery = list(golden["cmp_gallery"])
    probes = [golden[f"cmp_probe{q}"] for q in range(int(golden["cmp_q"]))]
    pairs = [int(x) for x in golden["cmp_pairs"]]
    rot = [float(x) for x in golden[f"cmp_{mode}_rot"]] or None
    scl = [float(x) for x in golden[f"cmp_{mode}_scl"]] or None
    config = {"comparison": {"n_processes": 3, "rotations": rot, "scales": scl}}
    ranks = compare_maps(probes, gallery, pairs, config)
    assert ranks.dtype == np.int32 and ranks.shape == (len(probes),)
    out = capsys.readouterr()
    assert "Print 0 true match ranked" in out.out + out.err
    from src.shoeprint_image_retrieval import similarity

    _check(similarity.last_result["scores"].cpu().numpy(), golden[f"cmp_{mode}_scores"])
    for q in range(len(probes)):
        lo, hi = ocmp.rank_interval(golden[f"cmp_{mode}_scores"][q], pairs[q])
        assert lo <= ranks[q] <= hi


def test_get_similarity_and_normxcorr_helpers(eng, golden):
    from src.shoeprint_image_retrieval.similarity import get_similarity, normxcorr

    for i in range(int(golden["gs_count"])):
        want = float(golden[f"gs{i}_out"])
        got = float(get_similarity(golden[f"gs{i}_p"], golden[f"gs{i}_g"]))
        assert abs(got - max(want, 0.0)) <= REL_TOL * max(abs(want), 1e-3)
    for i in range(int(golden["nx_count"])):
        got = normxcorr(golden[f"nx{i}_t"], golden[f"nx{i}_g"])
        np.testing.assert_allclose(got, golden[f"nx{i}_out"], rtol=0, atol=5e-6)


def test_idempotent_and_gallery_permutation_invariant(eng):
    """Size-independent properties: repeated runs are bit-identical (atomic max is order free);
    permuting the gallery permutes the score columns exactly."""
    from src.shoeprint_image_retrieval import synth

    gal = synth.device_gallery(51, 40, 16, 30, 20)
    prb, _ = synth.device_probes(52, gal, 20)
    ps = eng.MapSet.from_device(prb)
    for precision in ("fp16x3", "fp16_refine"):
        a = eng.score_matrix(ps, eng.MapSet.from_device(gal), [7], None, precision)
        b = eng.score_matrix(ps, eng.MapSet.from_device(gal), [7], None, precision)
        assert torch.equal(a, b)
        perm = torch.randperm(40, device="cuda")
        c = eng.score_matrix(ps, eng.MapSet.from_device(gal[perm].contiguous()), [7], None, precision)
        assert torch.equal(c, a[:, perm])


def test_full_bench_size_properties(eng):
    """BASELINE configs[1] at full size (1,500 probes x 150 gallery x 13 variants, 176x50x19): too big for
    the CPU oracle, so size-independent properties are checked: every probe is a noisy copy of its
    true match and must rank it first with a score near 1; the default precision mode must agree
    with fp16x3 to 1e-4; the score of an (unrotated) probe against itself-as-gallery is exactly the
    max over variants, so adding variants can only raise scores (monotonicity)."""
    from src.shoeprint_image_retrieval import synth

    gal = synth.device_gallery(61, 150, 176, 50, 19)
    prb, pairs = synth.device_probes(62, gal, 1500)
    ps, gs = eng.MapSet.from_device(prb), eng.MapSet.from_device(gal)
    rot = [a for a in range(-30, 31, 5) if a != 0]
    full = eng.score_matrix(ps, gs, rot, None)
    gt, ge, tv, ti, ts = eng.rank_true_matches(full, pairs, 5)
    assert int((gt == 0).sum()) == 1500
    assert torch.equal(ti[:, 0].long(), pairs.long())
    assert float(ts.min()) > 0.8
    x3 = eng.score_matrix(ps, gs, rot, None, "fp16x3")
    err = ((full - x3).abs() / x3.abs().clamp_min(1e-3)).max().item()
    assert err < REL_TOL, err
    base = eng.score_matrix(ps, gs, None, None)
    assert bool((full >= base).all())


def test_gallery_chunking_does_not_change_scores(eng):
    """Large galleries are scored chunk by chunk (engine.score_matrix gallery_chunk_bytes); chunking
    must be invisible: bit-identical scores, caller's gallery order."""
    from src.shoeprint_image_retrieval import synth

    gal = synth.device_gallery(71, 37, 8, 24, 16)
    prb, _ = synth.device_probes(72, gal, 9)
    ps, gs = eng.MapSet.from_device(prb), eng.MapSet.from_device(gal)
    whole = eng.score_matrix(ps, gs, [5], [1.1])
    per_map = gal[0].numel() * 4
    pieces = eng.score_matrix(ps, gs, [5], [1.1], gallery_chunk_bytes=5 * per_map)
    assert torch.equal(whole, pieces)


@pytest.mark.parametrize("precision", ["fp16_fp8c", "fp16x3", "fp16_refine"])
def test_heavy_tailed_feature_maps_stay_within_tolerance(eng, precision):
    """Post-activation CNN features are heavy tailed; cube the synthetic maps so single cells reach
    ~100x the typical magnitude of their channel and check the split-precision modes still meet 1e-4
    (the fp8 correction operands only span 2^-9..448 around the channel peak)."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    gallery = [np.ascontiguousarray((g / 6.0) ** 3 * 5.0, dtype=np.float32) for g in synth.make_gallery(81, 5, 6, 50, 19)]
    probes, pairs = synth.make_probes(82, gallery, 4, min_frac=1.0, noise=0.05)
    ranks, scores, _ = eng.compare(probes, gallery, pairs, [-5, 5], None, precision=precision)
    _, want = ocmp.compare_maps_oracle(probes, gallery, pairs, [-5, 5], None)
    _check(scores.cpu().numpy(), want)
    assert list(ranks) == [1, 1, 1, 1]


def test_refinement_finds_the_exact_maximum(eng):
    """The default mode screens in fp16 and re-evaluates only the positions within the candidate margin of a pair's
    screened maximum (engine.TAU_REL / TAU_ABS).  A missed maximum would show up as a score BELOW the float32 CUDA-core
    evaluation of the whole surface: check 64,000 pairs (x 4 variants) for that, and that the work list stays short
    (about one position per pair and launch, not one per variant or per patch)."""
    from src.shoeprint_image_retrieval import synth

    gal = synth.device_gallery(91, 160, 24, 34, 21)
    prb, _ = synth.device_probes(92, gal, 400)
    ps, gs = eng.MapSet.from_device(prb), eng.MapSet.from_device(gal)
    rot = [-10, 5, 20]
    exact = eng.score_matrix(ps, gs, rot, None, "fp32_simt")
    eng.collect_refine_stats = True
    eng.refine_stats(reset=True)
    try:
        got = eng.score_matrix(ps, gs, rot, None, "fp16_refine")
        stats = eng.refine_stats(reset=True)
    finally:
        eng.collect_refine_stats = False
    err = ((got - exact).abs() / exact.abs().clamp_min(1e-3)).max().item()
    assert err < 5e-6, err
    pairs = 400 * 160
    assert pairs <= stats["positions"] < 3 * pairs, stats
    screened = eng.score_matrix(ps, gs, rot, None, "fp16x1")
    assert ((screened - exact).abs() / exact.abs().clamp_min(1e-3)).max().item() > err  # the screen alone is looser


def test_edge_cases_match_reference_vectors(eng, golden_edge):
    """Dead (all-zero) and constant channels on either side, ReLU-like maps with flat halves, anti-correlated probes --
    what real post-activation feature maps contain and the smooth synthetic fields avoid -- against vectors produced by
    the unmodified reference (tests/golden/make_golden.py edge_cases).  A dead channel contributes exactly 0 on both
    sides: the reference maps its 0/0 to 0 (similarity.py:69-70), the library packs an all-zero template (E == 0) and
    zeroes the window norm where D <= 1e-10 * sum(g^2) (csrc/sir_pack.cu window_rnorm_kernel)."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval.similarity import compare_maps, get_similarity, last_result

    ge = golden_edge
    for i in range(int(ge["eg_count"])):
        want = float(ge[f"eg{i}_out"])
        got = float(get_similarity(ge[f"eg{i}_p"], ge[f"eg{i}_g"]))
        assert abs(got - want) <= REL_TOL * max(abs(want), 1e-3), (i, got, want)
    gallery = list(ge["ecmp_gallery"])
    probes = [ge[f"ecmp_probe{q}"] for q in range(int(ge["ecmp_q"]))]
    pairs = [int(x) for x in ge["ecmp_pairs"]]
    rot = [float(x) for x in ge["ecmp_rot"]]
    for precision in ("fp16_refine", "fp16_fp8c", "fp16x3"):
        ranks = compare_maps(probes, gallery, pairs, {"comparison": {"n_processes": 1, "rotations": rot, "scales": None, "precision": precision}})
        _check(last_result["scores"].cpu().numpy(), ge["ecmp_scores"])
        for q in range(len(probes)):
            lo, hi = ocmp.rank_interval(ge["ecmp_scores"][q], pairs[q])
            assert lo <= ranks[q] <= hi


def test_get_similarity_returns_negative_maxima(eng):
    """A lone get_similarity call is not floored at 0 in the reference (similarity.py:106-108)."""
    from oracle import ncc
    from src.shoeprint_image_retrieval.similarity import get_similarity

    rng = np.random.default_rng(5)
    found = 0
    for _ in range(4000):  # 2x3 templates over 2x2 cropped maps: now and then every position correlates negatively
        p = rng.standard_normal((1, 6, 7)).astype(np.float32)
        g = rng.standard_normal((1, 6, 6)).astype(np.float32)
        want = ncc.get_similarity(p, g, method="direct")
        if want < -0.05:
            got = float(get_similarity(p, g))
            assert abs(got - want) <= 1e-5, (got, want)
            found += 1
            if found == 3:
                return
    assert found > 0, "no all-negative surface drawn"


@pytest.mark.parametrize("shape", [(80, 59, 21), (176, 50, 19)])
def test_reference_shapes_against_the_cpu_oracle(eng, shape):
    """3 probes x 4 gallery maps at the reference's full map shapes (FID-300 block 4, 800x300 block 6), every parity mode
    against the float64 CPU oracle itself -- not against another GPU evaluation."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    c, h, w = shape
    gallery = synth.make_gallery(101, 4, c, h, w)
    probes, pairs = synth.make_probes(102, gallery, 3, min_frac=0.85)
    _, want = ocmp.compare_maps_oracle(probes, gallery, pairs, [-5], None, method="fast")
    for precision in ("fp16_refine", "fp16_fp8c", "fp16x3"):
        ranks, scores, _ = eng.compare(probes, gallery, pairs, [-5], None, precision=precision)
        try:
            _check(scores.cpu().numpy(), want)
        except AssertionError as exc:
            raise AssertionError(f"{precision}: {exc}") from None
        for i in range(len(probes)):
            lo, hi = ocmp.rank_interval(want[i], pairs[i])
            assert lo <= ranks[i] <= hi, precision


@pytest.mark.gpu
@pytest.mark.parametrize("byte", [0xFF, 0x7F, 0x7C])
def test_stale_shared_memory_does_not_leak_into_scores(eng, byte):
    """Shared memory keeps what the previous kernel left there.  An MMA over a partial K stage reads operand rows the kernel has
    not written in this pass (their partner taps are zero), so leftovers that decode to NaN / Inf (0xFF, 0x7F as e4m3 and
    float, 0x7C7C = +Inf as fp16) used to poison a position now and then.  Fill every SM's shared memory with such a
    pattern right before each scoring call (ragged probes: the multi-shape bucket path) and compare with the oracle."""
    import ctypes as C
    import torch
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import _native as nat, synth

    gallery = synth.make_gallery(101, 4, 80, 59, 21)
    probes, pairs = synth.make_probes(102, gallery, 3, min_frac=0.85)
    _, want = ocmp.compare_maps_oracle(probes, gallery, pairs, [-5], None, method="fast")
    uniform = [g[:, 2:-2, 1:-1].copy() for g in gallery[:3]]
    _, want_u = ocmp.compare_maps_oracle(uniform, gallery, [0, 1, 2], [-5], None, method="fast")
    real = nat.lib

    def poisoned(name):
        fn = getattr(real, name)

        def call(*args):
            nat.check(real.sir_debug_fill_shared_memory(byte, args[-1]), "sir_debug_fill_shared_memory")
            return fn(*args)

        return call

    class _Lib:  # the engine's kernels, each preceded by the fill on the same stream
        def __getattr__(self, name):
            if name in ("sir_ncc_screen", "sir_ncc_refine", "sir_ncc_scores", "sir_ncc_scores_multi", "sir_ncc_scores_fp8c"):
                return poisoned(name)
            return getattr(real, name)

    nat.lib = _Lib()
    try:
        for precision in ("fp16_refine", "fp16_fp8c", "fp16x3"):
            for p, w, pr in ((probes, want, pairs), (uniform, want_u, [0, 1, 2])):
                _, scores, _ = eng.compare(p, gallery, pr, [-5], None, precision=precision)
                try:
                    _check(scores.cpu().numpy(), w)
                except AssertionError as exc:
                    raise AssertionError(f"{precision}, fill {byte:#x}: {exc}") from None
    finally:
        nat.lib = real


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp16_refine", "fp16_fp8c", "fp16x3"])
def test_multi_tile_buckets_with_per_tile_row_ranges(eng, precision, monkeypatch):
    """Ragged probes whose merged shape buckets span several 256-column tiles: the columns are laid out tallest template
    first and the correlation kernel is told per tile which rows of the bucket layout are occupied (d_tile_rows), so tiles of
    short templates skip the rest.  Scores must equal the per-shape float32 CUDA-core evaluation (no buckets, no skipping),
    and a 3 x 4 subsample the float64 CPU oracle."""
    from oracle import compare as ocmp
    from src.shoeprint_image_retrieval import synth

    gallery = synth.make_gallery(201, 4, 12, 59, 21)
    probes, pairs = synth.make_probes(202, gallery, 100, min_frac=0.4)
    rotations = [r for r in range(-12, 13, 2) if r]
    seen = []
    real = eng._score_one_bucket

    def spy(members, gops, g0, scores, mode, flip, dev, approx=None):
        heights = {(hw[1] if flip else hw[0]) for hw, _ in members}
        seen.append((sum(eng._pad_cols(blk.ncols) for _, blk in members), len(heights)))
        return real(members, gops, g0, scores, mode, flip, dev, approx)

    monkeypatch.setattr(eng, "_score_one_bucket", spy)
    _, got, _ = eng.compare(probes, gallery, pairs, rotations, None, precision=precision)
    assert any(cols > 256 and nh > 4 for cols, nh in seen), f"no multi-tile bucket of mixed heights was launched: {seen}"
    monkeypatch.setattr(eng, "_score_one_bucket", real)
    _, want, _ = eng.compare(probes, gallery, pairs, rotations, None, precision="fp32_simt")
    try:
        _check(got.cpu().numpy(), want.cpu().numpy(), 2e-5 if precision == "fp16_refine" else REL_TOL)
        _, exact = ocmp.compare_maps_oracle(probes[:3], gallery, pairs[:3], rotations, None, method="fast")
        _check(got.cpu().numpy()[:3], exact)
    except AssertionError as exc:
        raise AssertionError(f"{precision}: {exc}") from None


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp16_refine", "fp16_fp8c", "fp16x3"])
def test_uninitialised_device_buffers_do_not_leak_into_scores(eng, precision, monkeypatch):
    """``torch.empty`` hands out whatever the allocator has; every device buffer the path allocates that way is filled
    with 0xFF (NaN in every format) here before use.  Scores of a uniform and of a ragged probe set (single-shape blocks
    and multi-shape buckets, rotations and scales) must be bit-identical to a normal run."""
    import torch
    from src.shoeprint_image_retrieval import synth

    gallery = synth.make_gallery(301, 5, 8, 40, 17)
    uniform, pairs_u = synth.make_probes(302, gallery, 24)
    ragged, pairs_r = synth.make_probes(303, gallery, 24, min_frac=0.5)
    cases = [(uniform, pairs_u, [-6, 6], None), (ragged, pairs_r, [-6, 6], [1.08]), (ragged, pairs_r, None, None)]
    want = [eng.compare(p, gallery, pr, rot, scl, precision=precision)[1].cpu().numpy() for p, pr, rot, scl in cases]
    real_empty = torch.empty

    def poisoned_empty(*args, **kwargs):
        t = real_empty(*args, **kwargs)
        if t.is_cuda and t.numel():
            t.view(torch.uint8).fill_(0xFF) if t.is_contiguous() else None
        return t

    monkeypatch.setattr(torch, "empty", poisoned_empty)
    got = [eng.compare(p, gallery, pr, rot, scl, precision=precision)[1].cpu().numpy() for p, pr, rot, scl in cases]
    monkeypatch.setattr(torch, "empty", real_empty)
    for g, w in zip(got, want):
        assert np.isfinite(g).all()
        np.testing.assert_array_equal(g, w)
