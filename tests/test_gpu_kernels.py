"""GPU parity tests of the individual libsir kernels against the CPU oracle / Pillow."""

import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as ge

    ge.build()
    from src.shoeprint_image_retrieval import engine

    assert torch.cuda.is_available()
    return engine


def test_device_is_sm100(eng):
    from src.shoeprint_image_retrieval import _native as nat

    sms, major, _ = nat.device_info()
    assert major == 10 and sms >= 100


def test_gallery_pack_and_window_rnorm(eng):
    from oracle import ncc

    rng = np.random.default_rng(0)
    maps = (rng.standard_normal((5, 3, 17, 12)) * 7 + 2).astype(np.float32)
    ms = eng.MapSet.from_host(list(maps))
    ops = eng.GalleryOperands.pack(ms.groups[0], keep_fp32=True)
    crop = maps[:, :, 2:-2, 2:-2]
    z = crop - crop.mean(axis=(2, 3), keepdims=True, dtype=np.float64).astype(np.float32)
    np.testing.assert_allclose(ops.gz.cpu().numpy(), z, rtol=0, atol=2e-6)
    e = ops.gexp.cpu().numpy()
    packed = (ops.ghi.float().cpu().numpy().astype(np.float64) + ops.glo.float().cpu().numpy())[..., : maps.shape[3] - 4]
    scaled = ops.gz.cpu().numpy().astype(np.float64) * (2.0 ** e)[:, :, None, None]
    assert np.abs(scaled).max() < 1024 and np.abs(scaled).reshape(5, 3, -1).max(-1).min() >= 512
    np.testing.assert_allclose(packed, scaled, rtol=0, atol=1024 * 2.0**-21)
    for hm, wm in [(13, 8), (5, 3), (20, 15), (2, 1)]:
        rn = ops.rnorm(hm, wm, simt=True).cpu().numpy().reshape(5, 3, 13, 8)
        for g in range(5):
            for c in range(3):
                d = ncc.window_denominator(ops.gz[g, c].cpu().numpy(), hm, wm)
                want = np.where(d > 0, 1 / np.sqrt(np.where(d > 0, d, 1)), 0)
                np.testing.assert_allclose(rn[g, c], want, rtol=2e-6, atol=0)


def test_rotate_bit_exact(eng, golden):
    from oracle import variants as ov

    rng = np.random.default_rng(1)
    for h, w in [(13, 9), (12, 12), (59, 21), (50, 19), (8, 30)]:
        m = rng.standard_normal((3, 2, h, w)).astype(np.float32)
        d = torch.from_numpy(m).cuda()
        for ang in [-30, -15, -9, -3, 3, 9, 15, 25, 45, 90, 180, 270, 360, 7.5, -180]:
            got = eng.make_variant(d, ang, None).cpu().numpy()
            want = np.stack([ov.rotate_maps(x, ang) for x in m])
            np.testing.assert_array_equal(got, want)
    angles = golden["var_rot_angles"]
    for i in range(int(golden["var_count"])):
        d = torch.from_numpy(golden[f"var{i}_in"][None]).cuda()
        for j, a in enumerate(angles):
            np.testing.assert_array_equal(eng.make_variant(d, float(a), None).cpu().numpy()[0], golden[f"var{i}_rot{j}"])


def test_resize_bit_exact(eng, golden):
    from oracle import variants as ov

    rng = np.random.default_rng(2)
    for h, w in [(13, 9), (59, 21), (50, 19), (9, 31)]:
        m = rng.standard_normal((2, 3, h, w)).astype(np.float32)
        d = torch.from_numpy(m).cuda()
        for s in [1.02, 1.04, 1.08, 0.9, 0.6, 1.3, 2.1]:
            got = eng.make_variant(d, None, s).cpu().numpy()
            want = np.stack([ov.resize_maps(x, s) for x in m])
            assert got.shape == want.shape
            np.testing.assert_array_equal(got, want)
    scales = golden["var_scales"]
    for i in range(int(golden["var_count"])):
        d = torch.from_numpy(golden[f"var{i}_in"][None]).cuda()
        for j, s in enumerate(scales):
            np.testing.assert_array_equal(eng.make_variant(d, None, float(s)).cpu().numpy()[0], golden[f"var{i}_scl{j}"])
    # rotate then resize, the order the reference applies them
    m = rng.standard_normal((1, 2, 21, 15)).astype(np.float32)
    got = eng.make_variant(torch.from_numpy(m).cuda(), -9.0, 1.08).cpu().numpy()[0]
    np.testing.assert_array_equal(got, ov.resize_maps(ov.rotate_maps(m[0], -9.0), 1.08))


def test_template_pack_layout(eng):
    from src.shoeprint_image_retrieval import _native as nat

    rng = np.random.default_rng(3)
    n, c, h, w = 3, 2, 11, 15  # Hm=7, Wm=11 -> 2 chunks per row
    m = rng.standard_normal((n, c, h, w)).astype(np.float32)
    m[1, 0] = 4.0  # flat channel: E == 0 -> all-zero template
    d = torch.from_numpy(m).cuda()
    hm, wm = h - 4, w - 4
    kpad = nat.lib.sir_template_kpad(hm, wm)
    ncols = 5
    thi = torch.full((c, ncols, kpad), 7.0, dtype=torch.float16, device="cuda")
    tlo = torch.full_like(thi, 7.0)
    t32 = torch.zeros((c, ncols, hm * wm), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.sir_template_pack(C.c_void_p(d.data_ptr()), n, c, h, w, 1, ncols, C.c_void_p(thi.data_ptr()),
                                        C.c_void_p(tlo.data_ptr()), C.c_void_p(t32.data_ptr()), st))
    crop = m[:, :, 2:-2, 2:-2].astype(np.float64)
    z = crop - crop.mean(axis=(2, 3), keepdims=True)
    e = (z * z).sum(axis=(2, 3), keepdims=True)
    tn = np.where(e > 1e-20, z / np.sqrt(np.where(e > 1e-20, e, 1)), 0)
    got32 = t32.cpu().numpy().reshape(c, ncols, hm, wm)
    full = (thi.double() + tlo.double()).cpu().numpy()
    for i in range(n):
        for ch in range(c):
            np.testing.assert_allclose(got32[ch, 1 + i], tn[i, ch], rtol=0, atol=3e-7)
            row = full[ch, 1 + i]
            body = row[: hm * 16].reshape(hm, 16)
            np.testing.assert_allclose(body[:, :wm], tn[i, ch] * 1024, rtol=0, atol=1024 * 2.0**-21)
            assert np.all(body[:, wm:] == 0) and np.all(row[hm * 16:] == 0)
    assert np.all(full[:, 0] == 7.0 * 2) and np.all(full[:, 4] == 14.0)  # untouched columns


def test_rank_topk_matches_numpy(eng):
    rng = np.random.default_rng(4)
    for q, g, k in [(7, 1000, 16), (3, 37, 8), (5, 4099, 64), (2, 5, 8), (4, 513, 0)]:
        s = rng.random((q, g)).astype(np.float32)
        s[:, ::7] = s[:, 3:4]  # exact ties
        true = rng.integers(0, g, size=q)
        d = torch.from_numpy(s).cuda()
        gt, ge, tv, ti, ts = eng.rank_true_matches(d, true, k, g0=100)
        ts_np = s[np.arange(q), true]
        # true index is global: shift by g0
        gt2, ge2, tv2, ti2, _ = eng.rank_true_matches(d, true + 100, k, g0=100)
        np.testing.assert_array_equal(gt2.cpu().numpy(), (s > ts_np[:, None]).sum(1))
        np.testing.assert_array_equal(ge2.cpu().numpy(), (s >= ts_np[:, None]).sum(1))
        if k:
            order = np.lexsort((np.arange(g)[None, :].repeat(q, 0), -s), axis=1)[:, :k]
            kk = min(k, g)
            np.testing.assert_array_equal(ti2.cpu().numpy()[:, :kk], order[:, :kk] + 100)
            np.testing.assert_array_equal(tv2.cpu().numpy()[:, :kk], np.take_along_axis(s, order[:, :kk], 1))
            if g < k:
                assert np.all(ti2.cpu().numpy()[:, g:] == -1)


def test_rank_topk_long_rows(eng):
    """Rows long enough for the multi-chunk fast path of the threshold select (sir_rank.cu), including the rows it has to
    redo: ascending values (every value beats the threshold: the candidate buffer overflows), descending, constant, and
    blocks of exact ties."""
    rng = np.random.default_rng(44)
    g = 50000
    rows = [rng.random(g), np.sort(rng.random(g)), np.sort(rng.random(g))[::-1], np.full(g, 0.25), np.repeat(rng.random(g // 100), 100),
            np.concatenate([rng.random(g // 2), np.sort(rng.random(g - g // 2))])]
    s = np.stack(rows).astype(np.float32)
    true = rng.integers(0, g, size=len(rows))
    d = torch.from_numpy(s).cuda()
    for k in (64, 128, 5):
        gt, ge, tv, ti, _ = eng.rank_true_matches(d, true, k, g0=7)
        ts_np = s[np.arange(len(rows)), true - 0]
        gt2, ge2, tv2, ti2, _ = eng.rank_true_matches(d, true + 7, k, g0=7)
        np.testing.assert_array_equal(gt2.cpu().numpy(), (s > ts_np[:, None]).sum(1))
        np.testing.assert_array_equal(ge2.cpu().numpy(), (s >= ts_np[:, None]).sum(1))
        order = np.lexsort((np.arange(g)[None, :].repeat(len(rows), 0), -s), axis=1)[:, :k]
        np.testing.assert_array_equal(ti2.cpu().numpy(), order + 7)
        np.testing.assert_array_equal(tv2.cpu().numpy(), np.take_along_axis(s, order, 1))


def test_merge_topk(eng):
    from src.shoeprint_image_retrieval import _native as nat

    rng = np.random.default_rng(5)
    p, q, k = 4, 6, 8
    vals = np.sort(rng.random((p, q, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    idx = rng.permutation(p * q * k).reshape(p, q, k).astype(np.int32)
    vals[3, :, 5:] = -np.inf
    idx[3, :, 5:] = -1
    dv, di = torch.from_numpy(vals).cuda(), torch.from_numpy(idx).cuda()
    ov = torch.empty((q, k), dtype=torch.float32, device="cuda")
    oi = torch.empty((q, k), dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.sir_merge_topk(C.c_void_p(dv.data_ptr()), C.c_void_p(di.data_ptr()), p, q, k, C.c_void_p(ov.data_ptr()), C.c_void_p(oi.data_ptr()), st))
    for qq in range(q):
        cand = sorted(((-vals[pp, qq, j], idx[pp, qq, j]) for pp in range(p) for j in range(k) if idx[pp, qq, j] >= 0))[:k]
        np.testing.assert_array_equal(oi.cpu().numpy()[qq], [c[1] for c in cand])
        np.testing.assert_array_equal(ov.cpu().numpy()[qq], [-c[0] for c in cand])


def test_device_lanczos_resize_bit_exact_vs_pillow(eng):
    """sir_image_resize_lanczos (the loader's resize on the device) against Pillow itself and the CPU restatement."""
    from PIL import Image

    from oracle.loader import resize_lanczos_u8

    rng = np.random.default_rng(9)
    images, sizes = [], []
    for t in range(24):
        h, w = (int(rng.integers(20, 200)), int(rng.integers(20, 120))) if t % 6 else (586, 270)
        s = float(rng.uniform(0.4, 1.5))
        images.append(rng.integers(0, 256, size=(h, w) if t % 4 else (h, w, 3), dtype=np.uint8))
        sizes.append((max(1, int(h * s)) if t % 5 else h, max(1, int(w * s))))
    images += [images[1].copy(), images[1].copy()]  # a batch of three with equal sizes
    sizes += [sizes[1], sizes[1]]
    got = eng.resize_images_lanczos(images, sizes)
    for im, (h2, w2), g in zip(images, sizes, got):
        want = np.array(Image.fromarray(im).resize((w2, h2), Image.Resampling.LANCZOS))
        np.testing.assert_array_equal(g, want)
        np.testing.assert_array_equal(resize_lanczos_u8(im, h2, w2), want)


def test_loader_with_device_resize_equals_host_loader(eng, tmp_path, monkeypatch):
    """Dataloader with SIR_DEVICE_RESIZE=1 (LANCZOS on the GPU) returns exactly what the host loader returns, on a directory
    whose prints are large enough to be scaled down (dataloader.py:408-417: scale = maximum_dim / largest)."""
    import sys

    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import synth_dataset

    from src.shoeprint_image_retrieval.dataloader import Dataloader

    root = tmp_path / "big"
    synth_dataset.write_dataset(root, 21, [(1800, 800)] * 4, [(1200 + 40 * i, 560 + 10 * i) for i in range(5)])
    cfg = synth_dataset.config_for(root, 1)
    monkeypatch.delenv("SIR_DEVICE_RESIZE", raising=False)
    host = list(Dataloader(cfg))
    monkeypatch.setenv("SIR_DEVICE_RESIZE", "1")
    dev = list(Dataloader(cfg))
    assert len(host) == len(dev) == 1
    assert Dataloader(cfg).scales[0] < 1.0
    for (m0, p0, pairs0, b0), (m1, p1, pairs1, b1) in zip(host, dev):
        assert pairs0 == pairs1 and b0 == b1
        for a, b in zip(m0 + p0, m1 + p1):
            np.testing.assert_array_equal(a, b)
