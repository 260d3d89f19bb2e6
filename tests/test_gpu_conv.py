"""GPU parity of the convolution entry points of the feature stage, operator by operator, against
torch.nn.functional in float64 (the torch reference of a floating-point kernel; tolerance: relative L2
<= 1e-5 per operator, i.e. float32-grade -- what float32 accumulation over up to ~1000 terms gives (measured
2e-7 ... 4e-6); a single fp16 product would be at ~3e-4).  Calls go through the
C ABI exactly as network.py issues them: split -> plan -> pack -> sir_feat_conv."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
F = torch.nn.functional

OP_REL_L2 = 1e-5


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _split_fp16(x):
    amax = float(x.abs().max())
    e = 0 if amax == 0 else 10 - int(np.frexp(amax)[1])
    xs = torch.ldexp(x.float(), torch.tensor(e))
    hi = xs.half()
    return hi.contiguous(), (xs - hi.float()).half().contiguous(), e


def _weights(nat, w, bk):
    """[cout][cin][kh][kw] float32 -> (whi, wlo, w_exp, rows, Kp) in the K order sir_feat_conv expects."""
    cout, cin, kh, kw = w.shape
    bn = int(nat.lib.sir_feat_conv_tile_n(cout))
    rows = (cout + bn - 1) // bn * bn
    cp = (cin + bk - 1) // bk * bk
    wm = torch.zeros((rows, kh * kw, cp))
    wm[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    hi, lo, e = _split_fp16(wm.reshape(rows, kh * kw * cp))
    return hi.cuda(), lo.cuda(), e, rows, kh * kw * cp


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _run_conv(nat, x, w, bias, pad, act, *, residual=None, chan_scale=None, planes_out=False, stride=1):
    """x [B,H,W,C] float32 cuda NHWC; returns (out [B,Ho,Wo,N] float32, (hi, lo, e) or None)."""
    b, h, wd, c = x.shape
    cout, cin, kh, kw = w.shape
    bk = 32 if cin % 32 == 0 else 16 if cin % 16 == 0 else 32
    st = _stream()
    amax = x.abs().max().reshape(1).float()
    xhi = torch.empty((b, h, wd, c), dtype=torch.float16, device="cuda")
    xlo = torch.empty_like(xhi)
    nat.check(nat.lib.sir_feat_im2col_split(_p(x), _p(amax), b, h, wd, c, 1, 1, 1, 0, None, c, _p(xhi), _p(xlo), st))
    whi, wlo, w_exp, rows, kp = _weights(nat, w, bk)
    per_image = 1 if chan_scale is not None else 0
    tile_n, granule, nbytes = C.c_int(), C.c_int(), C.c_longlong()
    nat.check(nat.lib.sir_feat_conv_plan(b, h, wd, c, kh, kw, pad, stride, bk, cout, per_image, C.byref(tile_n), C.byref(granule), C.byref(nbytes)))
    if per_image:
        pack = torch.empty((b, nbytes.value), dtype=torch.uint8, device="cuda")
        nat.check(nat.lib.sir_feat_conv_scale_weights(_p(whi), _p(wlo), rows, kp, kp // (kh * kw), cin, _p(chan_scale), b, tile_n.value,
                                                      granule.value, _p(pack), st))
    else:
        pack = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        nat.check(nat.lib.sir_feat_conv_pack_weights(_p(whi), _p(wlo), rows, kp, tile_n.value, granule.value, _p(pack), st))
    ho, wo = (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
    out = torch.empty((b, ho, wo, cout), dtype=torch.float32, device="cuda")
    amax_out = torch.zeros(1, device="cuda")
    bias_d = bias.float().cuda().contiguous()
    ohi = olo = exp_out = amax_res = None
    if planes_out:
        ohi = torch.empty((b, ho, wo, cout), dtype=torch.float16, device="cuda")
        olo = torch.empty_like(ohi)
        exp_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    if residual is not None:
        amax_res = residual.abs().max().reshape(1).float()
    bound_mult = float(w.double().abs().flatten(1).sum(1).max()) * (1 + 1e-5)
    bound_add = float(bias.abs().max()) * (1 + 1e-5)
    nat.check(nat.lib.sir_feat_conv(_p(xhi), _p(xlo), _p(amax), b, h, wd, c, kh, kw, pad, stride, bk, _p(pack), tile_n.value, granule.value, per_image,
                                    cout, w_exp, _p(bias_d), _p(residual), act, _p(out), cout, _p(amax_out), None, _p(ohi), _p(olo),
                                    _p(exp_out), bound_mult, bound_add, _p(amax_res), st))
    torch.cuda.synchronize()
    assert abs(float(amax_out) - float(out.abs().max())) <= 1e-6 * float(out.abs().max())
    return out, ((ohi, olo, int(exp_out)) if planes_out else None)


def _reference(x, w, bias, pad, act, residual=None, chan_scale=None, stride=1):
    xd = x.double().permute(0, 3, 1, 2)
    if chan_scale is not None:
        xd = xd * chan_scale.double()[:, :, None, None]
    y = F.conv2d(xd, w.double().cuda(), bias.double().cuda(), padding=pad, stride=stride)
    y = F.silu(y) if act == 1 else F.relu(y) if act == 2 else y
    y = y.permute(0, 2, 3, 1)
    return y + residual.double() if residual is not None else y


CASES = [
    # (B, H, W, Cin, Cout, k, pad, act)
    (2, 37, 29, 24, 24, 3, 1, 1),     # stage-1 FusedMBConv shape: C padded to the K step, narrow N
    (3, 40, 21, 48, 192, 3, 1, 1),    # halo kernel, two 16-channel K chunks + a half chunk
    (2, 19, 50, 80, 320, 3, 1, 0),    # two N tiles
    (1, 9, 7, 64, 64, 3, 1, 2),       # image smaller than one patch, ReLU (VGG)
    (2, 25, 19, 176, 1056, 1, 0, 1),  # 1x1 expansion, six N tiles, ragged last row tile
    (2, 25, 19, 1056, 176, 1, 0, 0),  # 1x1 projection, long K
    (1, 64, 48, 32, 40, 5, 2, 0),     # 5x5 kernel, N not a multiple of 16
]


@pytest.mark.parametrize("b,h,w,cin,cout,k,pad,act", CASES)
def test_conv_matches_torch(b, h, w, cin, cout, k, pad, act):
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(1000 * cin + cout + k)
    x = (torch.randn((b, h, w, cin), generator=g) * 3).cuda()
    wt = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    out, _ = _run_conv(nat, x, wt, bias, pad, act)
    rel = _rel(out, _reference(x, wt, bias, pad, act))
    assert rel < OP_REL_L2, rel


@pytest.mark.parametrize("halo", ["0", "1"])
def test_conv_patch_and_halo_kernels_agree(halo, monkeypatch):
    """Both kernels of sir_feat_conv (per-tap patch loads / one halo load) on the same 3x3 layer, with residual."""
    import subprocess
    import sys
    from pathlib import Path

    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import importlib.util, torch\n"
        "spec = importlib.util.spec_from_file_location('t', %r); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)\n"
        "from src.shoeprint_image_retrieval import _native as nat\n"
        "g = torch.Generator().manual_seed(5)\n"
        "x = torch.randn((2, 33, 26, 48), generator=g).cuda(); w = torch.randn((96, 48, 3, 3), generator=g) / 20; b = torch.randn(96, generator=g)\n"
        "r = torch.randn((2, 33, 26, 96), generator=g).cuda()\n"
        "out, _ = t._run_conv(nat, x, w, b, 1, 1, residual=r)\n"
        "print('REL', t._rel(out, t._reference(x, w, b, 1, 1, residual=r)))\n"
    ) % (str(Path(__file__).resolve().parents[1]), str(Path(__file__).resolve().parents[1]), str(Path(__file__).resolve()))
    env = dict(__import__("os").environ, SIR_CONV_HALO=halo)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    rel = float(res.stdout.split("REL")[1])
    assert rel < OP_REL_L2


def test_conv_operand_planes_chain():
    """Epilogue-written operand planes: hi + lo reproduce the float32 output to 2^-20 of the a-priori bound, and a second
    convolution fed from them (d_exp_in) matches torch on the composed function."""
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 30, 22, 48), generator=g).cuda()
    w1 = torch.randn((192, 48, 3, 3), generator=g) / 20
    b1 = torch.randn(192, generator=g) * 0.1
    out1, (hi, lo, e) = _run_conv(nat, x, w1, b1, 1, 1, planes_out=True)
    recon = torch.ldexp(hi.float() + lo.float(), torch.tensor(-e, device="cuda"))
    assert float((recon - out1).abs().max()) <= float(out1.abs().max()) * 2.0 ** -18
    assert float(hi.float().abs().max()) < 2.0 ** 15  # the bound kept fp16 far from overflow

    w2 = torch.randn((48, 192, 1, 1), generator=g) / 14
    b2 = torch.randn(48, generator=g) * 0.1
    st = _stream()
    bk = 32
    whi, wlo, w_exp, rows, kp = _weights(nat, w2, bk)
    bsz, h, wd, c = out1.shape
    tile_n, granule, nbytes = C.c_int(), C.c_int(), C.c_longlong()
    nat.check(nat.lib.sir_feat_conv_plan(bsz, h, wd, c, 1, 1, 0, 1, bk, 48, 0, C.byref(tile_n), C.byref(granule), C.byref(nbytes)))
    pack = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    nat.check(nat.lib.sir_feat_conv_pack_weights(_p(whi), _p(wlo), rows, kp, tile_n.value, granule.value, _p(pack), st))
    out2 = torch.empty((bsz, h, wd, 48), dtype=torch.float32, device="cuda")
    amax1 = out1.abs().max().reshape(1)
    e_dev = torch.tensor([e], dtype=torch.int32, device="cuda")
    nat.check(nat.lib.sir_feat_conv(_p(hi), _p(lo), _p(amax1), bsz, h, wd, c, 1, 1, 0, 1, bk, _p(pack), tile_n.value, granule.value, 0, 48, w_exp,
                                    _p(b2.cuda()), None, 0, _p(out2), 48, None, _p(e_dev), None, None, None, 0.0, 0.0, None, st))
    torch.cuda.synchronize()
    ref = _reference(_reference(x, w1, b1, 1, 1).float(), w2, b2, 0, 0)
    assert _rel(out2, ref) < 2 * OP_REL_L2


@pytest.mark.parametrize("b,h,w,cin,cout,k,pad,stride", [(2, 41, 30, 24, 96, 3, 1, 2), (1, 64, 37, 48, 192, 3, 1, 2), (2, 33, 20, 64, 128, 1, 0, 2), (1, 50, 31, 32, 64, 5, 2, 3)])
def test_strided_conv_matches_torch(b, h, w, cin, cout, k, pad, stride):
    """Strided convolutions: the patch of every tap is fetched with TMA element strides (no im2col matrix)."""
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(77 + cin + stride)
    x = (torch.randn((b, h, w, cin), generator=g) * 2).cuda()
    wt = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    out, _ = _run_conv(nat, x, wt, bias, pad, 1, stride=stride)
    rel = _rel(out, _reference(x, wt, bias, pad, 1, stride=stride))
    assert rel < OP_REL_L2, rel


def test_conv_per_image_scaled_weights():
    """SqueezeExcitation folded into the projection: W_b = W * scale[b] per image equals scaling the input channels."""
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(12)
    x = torch.randn((3, 25, 19, 320), generator=g).cuda()
    w = torch.randn((160, 320, 1, 1), generator=g) / 18
    b = torch.randn(160, generator=g) * 0.1
    scale = torch.rand((3, 320), generator=g).cuda()
    r = torch.randn((3, 25, 19, 160), generator=g).cuda()
    out, _ = _run_conv(nat, x, w, b, 0, 0, residual=r, chan_scale=scale)
    assert _rel(out, _reference(x, w, b, 0, 0, residual=r, chan_scale=scale)) < OP_REL_L2


@pytest.mark.parametrize("stride,act", [(1, 1), (2, 1), (1, 0)])
def test_depthwise_with_planes_and_pool(stride, act):
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(13 + stride)
    bsz, h, wd, c = 2, 27, 21, 136
    x = torch.randn((bsz, h, wd, c), generator=g).cuda()
    w = torch.randn((c, 1, 3, 3), generator=g) / 3
    bias = torch.randn(c, generator=g) * 0.2
    ho, wo = (h + 2 - 3) // stride + 1, (wd + 2 - 3) // stride + 1
    out = torch.empty((bsz, ho, wo, c), device="cuda")
    hi = torch.empty((bsz, ho, wo, c), dtype=torch.float16, device="cuda")
    lo = torch.empty_like(hi)
    e = torch.zeros(1, dtype=torch.int32, device="cuda")
    parts = int(nat.lib.sir_feat_dwconv_pool_parts(3, stride, c, ho, wo))
    part = torch.empty((bsz, parts, c), device="cuda")
    amax_in, amax_out = x.abs().max().reshape(1), torch.zeros(1, device="cuda")
    w_d = w[:, 0].permute(1, 2, 0).contiguous().cuda()
    nat.check(nat.lib.sir_feat_dwconv(_p(x), bsz, h, wd, c, 3, stride, 1, _p(w_d), _p(bias.cuda()), act, _p(out), _p(amax_out), _p(part), _p(amax_in),
                                      _p(hi), _p(lo), _p(e), float(w.abs().flatten(1).sum(1).max()) * 1.00001, float(bias.abs().max()) * 1.00001,
                                      _stream()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().cuda(), bias.double().cuda(), stride=stride, padding=1, groups=c)
    ref = (F.silu(ref) if act == 1 else ref).permute(0, 2, 3, 1)
    assert _rel(out, ref) < OP_REL_L2
    recon = torch.ldexp(hi.float() + lo.float(), torch.tensor(-int(e), device="cuda"))
    assert float((recon - out).abs().max()) <= float(out.abs().max()) * 2.0 ** -18
    assert _rel(part.sum(1), ref.sum((1, 2))) < 1e-5
    assert abs(float(amax_out) - float(out.abs().max())) <= 1e-6 * float(out.abs().max())


def test_stem_conv_c3k3():
    from src.shoeprint_image_retrieval import _native as nat

    g = torch.Generator().manual_seed(14)
    x = torch.randn((2, 61, 45, 3), generator=g).cuda()
    w = torch.randn((24, 3, 3, 3), generator=g) / 5
    bias = torch.randn(24, generator=g) * 0.2
    ho, wo = (61 + 2 - 3) // 2 + 1, (45 + 2 - 3) // 2 + 1
    out = torch.empty((2, ho, wo, 24), device="cuda")
    amax_in, amax_out = x.abs().max().reshape(1), torch.zeros(1, device="cuda")
    w_d = w.permute(0, 2, 3, 1).reshape(24, 27).t().contiguous().cuda()
    nat.check(nat.lib.sir_feat_conv_c3k3(_p(x), _p(amax_in), 2, 61, 45, 2, 1, _p(w_d), _p(bias.cuda()), 24, 1, _p(out), _p(amax_out), None, None, None,
                                         0.0, 0.0, _stream()))
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(x.double().permute(0, 3, 1, 2), w.double().cuda(), bias.double().cuda(), stride=2, padding=1)).permute(0, 2, 3, 1)
    assert _rel(out, ref) < OP_REL_L2


def test_conv_rejects_mismatched_pack():
    from src.shoeprint_image_retrieval import _native as nat

    x = torch.zeros((1, 8, 8, 32), dtype=torch.float16, device="cuda")
    amax = torch.ones(1, device="cuda")
    pack = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    out = torch.zeros((1, 8, 8, 64), device="cuda")
    bias = torch.zeros(64, device="cuda")
    rc = nat.lib.sir_feat_conv(_p(x), _p(x), _p(amax), 1, 8, 8, 32, 1, 1, 0, 1, 32, _p(pack), 16, 32, 0, 64, 0, _p(bias), None, 0, _p(out), 64, None,
                               None, None, None, None, 0.0, 0.0, None, _stream())
    assert rc != 0 and b"packed for" in nat.lib.sir_last_error()
