"""ctypes binding of ``libsir.so`` (C ABI in ``include/sir.h``).

There is no CPU fallback: importing this module without the built library raises, and every
entry point raises ``SirError`` with the library's message on a non-zero return code.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libsir.so"

PREC_FP16X3 = 0
PREC_FP16X1 = 1
PREC_FP32_SIMT = 2
PREC_FP16_FP8C = 3
PREC_FP16_REFINE = 4
PRECISIONS = {"fp16x3": PREC_FP16X3, "fp16x1": PREC_FP16X1, "fp32_simt": PREC_FP32_SIMT, "fp16_fp8c": PREC_FP16_FP8C,
              "fp16_refine": PREC_FP16_REFINE}

_p = C.c_void_p
_i = C.c_int

# name -> (restype, argtypes); mirrors include/sir.h one to one
SIGNATURES = {
    "sir_last_error": (C.c_char_p, []),
    "sir_abi_version": (_i, []),
    "sir_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "sir_gallery_pitch": (_i, [_i]),
    "sir_gallery_pack": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "sir_gallery_window_rnorm": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "sir_gallery_window_rnorm_multi": (_i, [_p, _p, _i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_p), _p]),
    "sir_variant_rotate": (_i, [_p, _i, _i, _i, _i, C.c_double, _p, _p]),
    "sir_variant_resize_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "sir_variant_resize": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, C.c_size_t, _p]),
    "sir_image_resize_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "sir_image_resize_lanczos": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, C.c_size_t, _p]),
    "sir_maps_transpose": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "sir_template_kpad": (_i, [_i, _i]),
    "sir_template_pack": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sir_ncc_scores": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p]),
    "sir_ncc_surface": (_i, [_p, _p, _i, _i, _p, _i, _i, _p, _p]),
    "sir_gallery_pitch8": (_i, [_i]),
    "sir_gallery_pack_fp8c": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "sir_template_kpad_fp8c": (_i, [_i, _i]),
    "sir_template_pack_fp8c": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sir_ncc_scores_fp8c": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p]),
    "sir_template_pack_embed": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "sir_ncc_scores_multi": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p]),
    "sir_ncc_screen_rec_count": (C.c_longlong, [_i, _i, _i, _i]),
    "sir_gallery_pack_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "sir_template_pack_screen": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sir_variant_index_map": (_i, [_i, _i, C.c_double, _i, _p, _p]),
    "sir_ncc_screen": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, C.c_float, C.c_float, _p, _p, _p]),
    "sir_ncc_refine": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, C.c_float, C.c_float, _p, _i, _p, _p]),
    "sir_memset_zero": (_i, [_p, C.c_size_t, _p]),
    "sir_ncc_norm_chunk": (_i, []),
    "sir_debug_fill_shared_memory": (_i, [_i, _p]),
    "sir_ncc_cost": (_i, [_i, _i, _i, _i, _i, _i, C.POINTER(C.c_double)]),
    "sir_true_scores": (_i, [_p, _i, _i, _i, _p, _i, _p, _p]),
    "sir_rank_topk": (_i, [_p, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p]),
    "sir_merge_topk": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "sir_scatter_columns": (_i, [_p, _i, _i, _i, _p, _p, _i, _p]),
    "sir_feat_clahe_to_nhwc": (_i, [_p, _i, _i, _i, C.c_double, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _p, _p, _p, _p]),
    "sir_feat_clahe_rgb_to_nhwc": (_i, [_p, _i, _i, _i, C.c_double, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _p, _p, _p, _p, _p, _p,
                                        _p, _p]),
    "sir_feat_image_to_nhwc": (_i, [_p, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _p, _p]),
    "sir_feat_im2col_split": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    "sir_feat_conv_tile_n": (_i, [_i]),
    "sir_feat_conv_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_longlong)]),
    "sir_feat_conv_scale_weights": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _p]),
    "sir_feat_conv_pack_weights": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "sir_feat_conv": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p,
                            _p, _p, _p, _p, C.c_float, C.c_float, _p, _p]),
    "sir_feat_conv_c3k3": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, C.c_float, C.c_float, _p]),
    "sir_feat_dwconv_pool_parts": (_i, [_i, _i, _i, _i, _i]),
    "sir_feat_dwconv": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, C.c_float, C.c_float, _p]),
    "sir_feat_pool_sum": (_i, [_p, _i, _i, _i, _p, _p]),
    "sir_feat_se_scale": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "sir_feat_maxpool": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "sir_feat_affine_act": (_i, [_p, C.c_longlong, _i, _i, _i, _p, _p, _i, _p, _p, _p]),
    "sir_feat_avgpool2d": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "sir_feat_nhwc_to_nchw": (_i, [_p, _i, _i, _i, _p, _p]),
}


class SirError(RuntimeError):
    """A libsir entry point returned a non-zero code."""


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        msg = (
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the matching path."
        )
        raise ImportError(msg)
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.sir_last_error().decode(errors="replace")
        raise SirError(f"{what or 'libsir'} failed ({rc}): {msg}")


def device_info() -> tuple[int, int, int]:
    sm, major, minor = _i(), _i(), _i()
    check(lib.sir_device_info(C.byref(sm), C.byref(major), C.byref(minor)), "sir_device_info")
    return sm.value, major.value, minor.value
