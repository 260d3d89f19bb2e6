"""GPU parity of the feature stage (network.Model through libsir) against the same torchvision
modules evaluated by PyTorch in float32 (TF32 off) -- the torch fp32 reference of a floating-point
kernel.  Metric: relative L2 error of the feature maps (the feature stage has its own
tolerance, 1e-4 relative L2, because 1e-4 on scores is defined on identical feature-map inputs)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FEATURE_REL_L2 = 1e-4


def _config(model_type):
    return {"model": {"type": model_type, "clahe_clip_limit": 2.0, "clahe_tile_grid_size": [8, 8]}}


def _randomise_bn(seq, seed):
    g = torch.Generator().manual_seed(seed)
    for m in seq.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def _torch_reference(model, img):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = model.transform_rgb(img) if img.ndim == 3 else model.transform(img)
    net = model.model.to("cuda").double()
    with torch.no_grad():
        y = net(x[None].to("cuda").double())
    model.model.to("cpu").float()
    return y[0].float().cpu().numpy()


def _image(seed, h, w, rgb=False):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h // 4 + 1, w // 4 + 1, 3 if rgb else 1)).astype(np.float32)
    img = np.kron(base, np.ones((4, 4, 1), np.float32))[:h, :w]
    img = np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)
    return img if rgb else img[..., 0]


def _rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("model_type,block,hw,rgb", [
    ("EfficientNetV2_M", 4, (150, 94), False),
    ("EfficientNetV2_M", 6, (131, 77), False),
    ("EfficientNetV2_S", 5, (96, 64), True),
    ("EfficientNet_B1", 5, (90, 70), False),
    ("VGG16", 9, (70, 50), False),
    ("VGG19_BN", 12, (64, 48), True),
    ("DenseNet_201", 5, (96, 64), False),
    ("DenseNet_201", 7, (128, 96), False),
    ("DenseNet_201", 12, (96, 96), True),
])
def test_feature_maps_match_torch_fp32(model_type, block, hw, rgb):
    import __graft_entry__ as ge

    ge.build()
    from src.shoeprint_image_retrieval import network

    model = network.Model(_config(model_type), block, random_init_seed=3)
    _randomise_bn(model.model, 5)
    model.program = network._Program(list(model.model.children()), model.device)
    img = _image(7, *hw, rgb=rgb)
    got = model.get_feature_maps(img)
    want = _torch_reference(model, model._clahe(img))
    assert got.shape == want.shape == tuple(network.get_output_size(model, (1, 3, *hw)))[1:]
    assert got.dtype == np.float32
    err = _rel_l2(got, want)
    assert err < FEATURE_REL_L2, f"{model_type}[:{block}] relative L2 error {err:.3e}"


def test_multiple_feature_maps_batches_and_keeps_order():
    from src.shoeprint_image_retrieval import network

    model = network.Model(_config("EfficientNetV2_M"), 4, random_init_seed=1)
    imgs = [_image(1, 80, 56), _image(2, 64, 48), _image(3, 80, 56), _image(4, 80, 56)]
    many = model.get_multiple_feature_maps(imgs, progress=False)
    for im, fm in zip(imgs, many):
        one = model.get_feature_maps(im)
        assert fm.shape == one.shape == (80, (im.shape[0] + 7) // 8, (im.shape[1] + 7) // 8)
        assert _rel_l2(fm, one) < 1e-6


def test_unknown_model_string_raises_lookup_error():
    from src.shoeprint_image_retrieval import network

    with pytest.raises(LookupError, match="Model string not found"):
        network.Model(_config("ResNet50"), 4, random_init_seed=0)


def test_gpu_clahe_bit_exact_vs_opencv():
    """sir_feat_clahe_to_nhwc against cv2.createCLAHE(...).apply: the equalised uint8 image must be
    identical, and the fused normalised output must equal ToTensor + repeat + Normalize of it."""
    import ctypes as C

    import cv2

    from src.shoeprint_image_retrieval import _native as nat

    rng = np.random.default_rng(9)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    for (h, w), clip, tiles in [((470, 162), 2.0, (8, 8)), ((800, 300), 2.0, (8, 8)), ((123, 77), 4.0, (4, 6)), ((64, 64), 0.0, (8, 8)), ((97, 45), 40.0, (2, 2))]:
        imgs = np.stack([_image(int(rng.integers(1 << 30)), h, w) for _ in range(3)])
        d = torch.from_numpy(imgs).cuda()
        lut = torch.empty((3, tiles[0] * tiles[1], 256), dtype=torch.uint8, device="cuda")
        u8 = torch.empty_like(d)
        out = torch.empty((3, h, w, 3), dtype=torch.float32, device="cuda")
        amax = torch.zeros(1, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib.sir_feat_clahe_to_nhwc(C.c_void_p(d.data_ptr()), 3, h, w, clip, tiles[0], tiles[1], (C.c_float * 3)(*mean), (C.c_float * 3)(*std),
                                                 C.c_void_p(lut.data_ptr()), C.c_void_p(u8.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(amax.data_ptr()), st))
        cl = cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles)
        want = np.stack([cl.apply(im) for im in imgs])
        np.testing.assert_array_equal(u8.cpu().numpy(), want, err_msg=f"{h}x{w} clip {clip} tiles {tiles}")
        x = torch.from_numpy(want).float().div(255)[..., None].repeat(1, 1, 1, 3)
        ref = (x - torch.tensor(mean)) / torch.tensor(std)
        np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())
        assert abs(float(amax) - float(ref.abs().max())) < 1e-6


def test_model_gpu_clahe_path_matches_host_clahe_path(monkeypatch):
    from src.shoeprint_image_retrieval import network

    img = _image(21, 300, 130)
    model = network.Model(_config("EfficientNetV2_M"), 4, random_init_seed=2)
    a = model.get_feature_maps(img)
    model._host_clahe = True
    b = model.get_feature_maps(img)
    np.testing.assert_array_equal(a, b)


def test_gpu_rgb_clahe_bit_exact_vs_opencv():
    """sir_feat_clahe_rgb_to_nhwc against the reference's RGB branch (network.py:199-204): cv2 RGB2LAB, CLAHE on L, LAB2RGB.
    The equalised uint8 image must be identical and the fused output equal to ToTensor + Normalize of it."""
    import ctypes as C

    import cv2

    from src.shoeprint_image_retrieval import _native as nat, network

    rng = np.random.default_rng(11)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    rgb2lab, lab2rgb = network._lab_tables(torch.device("cuda", 0))
    p = lambda t: C.c_void_p(t.data_ptr())
    for (h, w), clip, tiles in [((470, 162), 2.0, (8, 8)), ((123, 77), 4.0, (4, 6)), ((64, 64), 0.0, (8, 8))]:
        smooth = np.stack([np.stack([_image(int(rng.integers(1 << 30)), h, w) for _ in range(3)], -1) for _ in range(2)])
        noise = rng.integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)  # every corner of the colour cube
        imgs = np.concatenate([smooth, noise])
        n = len(imgs)
        d = torch.from_numpy(imgs).cuda()
        lut = torch.empty((n, tiles[0] * tiles[1], 256), dtype=torch.uint8, device="cuda")
        l_plane = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        ab_plane = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
        u8 = torch.empty_like(d)
        out = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
        amax = torch.zeros(1, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib.sir_feat_clahe_rgb_to_nhwc(p(d), n, h, w, clip, tiles[0], tiles[1], (C.c_float * 3)(*mean), (C.c_float * 3)(*std),
                                                     p(rgb2lab), p(lab2rgb), p(l_plane), p(ab_plane), p(lut), p(u8), p(out), p(amax), st))
        cl = cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles)
        want = []
        for im in imgs:
            l_ch, a_ch, b_ch = cv2.split(cv2.cvtColor(im, cv2.COLOR_RGB2LAB))
            want.append(cv2.cvtColor(cv2.merge((cl.apply(l_ch), a_ch, b_ch)), cv2.COLOR_LAB2RGB))
        want = np.stack(want)
        np.testing.assert_array_equal(u8.cpu().numpy(), want, err_msg=f"{h}x{w} clip {clip} tiles {tiles}")
        ref = (torch.from_numpy(want).float().div(255) - torch.tensor(mean)) / torch.tensor(std)
        np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())
        assert abs(float(amax) - float(ref.abs().max())) < 1e-6


def test_model_rgb_prints_gpu_clahe_matches_host_clahe():
    """RGB prints through the Model: CLAHE on the device (tabulated LAB round trip) and on the host (cv2) give identical maps,
    single image and batched."""
    from src.shoeprint_image_retrieval import network

    imgs = [np.stack([_image(31 + 3 * i + c, 300, 130) for c in range(3)], -1) for i in range(3)]
    model = network.Model(_config("EfficientNetV2_M"), 4, random_init_seed=2)
    a = model.get_feature_maps(imgs[0])
    many = model.get_multiple_feature_maps(imgs, progress=False)
    model._host_clahe = True
    network.clear_caches()
    b = model.get_feature_maps(imgs[0])
    many_host = model.get_multiple_feature_maps(imgs, progress=False)
    np.testing.assert_array_equal(a, b)
    for x, y in zip(many, many_host):
        np.testing.assert_array_equal(x, y)
    # a batch shares one running |max| for the operand scaling, so batched and single maps agree to rounding only
    assert np.linalg.norm(many[0] - a) <= 1e-5 * np.linalg.norm(a)


def test_feature_stage_ignores_stale_shared_memory():
    """Every feature-stage kernel after a 0xFF fill of all shared memory (NaN as fp16 / fp32), every ``torch.empty`` device
    buffer pre-filled with 0xFF: the maps must be bit-identical to a normal run, i.e. no MMA, halo or reduction reads a
    shared-memory or global cell the stage has not written itself."""
    from src.shoeprint_image_retrieval import _native as nat, network

    imgs = [_image(71 + i, 300, 130) for i in range(3)]
    rgb = [np.stack([_image(81 + c, 200, 100) for c in range(3)], -1)]
    for name, block in (("EfficientNetV2_M", 5), ("VGG16", 5)):
        model = network.Model(_config(name), block, random_init_seed=4)
        network.clear_caches()
        want = [m.copy() for m in model.get_multiple_feature_maps(imgs, progress=False)] + [model.get_feature_maps(rgb[0])]
        real = nat.lib

        class _Lib:
            def __getattr__(self, fn_name):
                fn = getattr(real, fn_name)
                if not fn_name.startswith("sir_feat_") or fn_name.endswith(("_plan", "_tile_n", "_parts", "_bytes")):
                    return fn

                def call(*args):
                    nat.check(real.sir_debug_fill_shared_memory(0xFF, args[-1]), "sir_debug_fill_shared_memory")
                    return fn(*args)

                return call

        real_empty = torch.empty

        def poisoned_empty(*args, **kwargs):  # and every device buffer the stage allocates starts as 0xFF bytes
            t = real_empty(*args, **kwargs)
            if t.is_cuda and t.numel() and t.is_contiguous():
                t.view(torch.uint8).fill_(0xFF)
            return t

        nat.lib = _Lib()
        torch.empty = poisoned_empty
        try:
            network.clear_caches()
            got = [m.copy() for m in model.get_multiple_feature_maps(imgs, progress=False)] + [model.get_feature_maps(rgb[0])]
        finally:
            nat.lib = real
            torch.empty = real_empty
        for a, b in zip(got, want):
            assert np.isfinite(a).all()
            np.testing.assert_array_equal(a, b, err_msg=name)


def test_device_resident_handoff_to_compare():
    """maps returned by get_multiple_feature_maps keep device copies; compare uses them (no H2D) and gives
    the same ranks and scores as the host lists.  Replacing an element falls back to the host path."""
    from src.shoeprint_image_retrieval import engine, network

    model = network.Model(_config("EfficientNetV2_M"), 5, random_init_seed=3)
    prints = [_image(100 + i, 192 + 32 * (i % 2), 128) for i in range(6)]
    marks = [p[16:-16, 8:-8].copy() for p in prints[:4]]
    g_maps = model.get_multiple_feature_maps(prints, progress=False)
    p_maps = model.get_multiple_feature_maps(marks, progress=False)
    assert isinstance(g_maps, list) and g_maps.device_copies() is not None
    for a, (t, idx) in [(g_maps, grp) for grp in g_maps.device_copies()]:
        for j, i in enumerate(idx):
            np.testing.assert_array_equal(a[i], t[j].cpu().numpy())
    pairs = np.arange(4)
    assert engine.MapSet.from_host(g_maps).h2d_bytes == 0
    assert engine.MapSet.from_host(list(g_maps)).h2d_bytes > 0
    r_dev, s_dev, _ = engine.compare(p_maps, g_maps, pairs, [-10, 10], None)
    r_host, s_host, _ = engine.compare(list(p_maps), list(g_maps), pairs, [-10, 10], None)
    np.testing.assert_array_equal(r_dev, r_host)
    assert torch.equal(s_dev, s_host)
    g_maps[0] = g_maps[0] * 1.0  # a different array object: the device copy is no longer trusted
    assert g_maps.device_copies() is None and engine.MapSet.from_host(g_maps).h2d_bytes > 0
