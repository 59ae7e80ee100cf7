"""ctypes binding of `include/diffmvs_b200.h` (the C ABI of the CUDA kernel library).

The product path has no CPU fallback: if the shared library is missing or fails to load,
`lib()` raises.  Build it with `python -m diffmvs_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdiffmvs_b200.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_SILU = 0, 1, 2, 3, 4
RES_NONE, RES_PRE_ACT, RES_POST_ACT = 0, 1, 2
EPI_STD, EPI_GRU_ZR, EPI_GRU_Q = 0, 1, 2
PREC_FP32, PREC_TF32X3, PREC_TF32, PREC_TC_TF32X3, PREC_TC_TF32, PREC_AUTO = 0, 1, 2, 3, 4, 5
PREC_WS_TF32X3, PREC_WS_TF32, PREC_WS2_TF32X3, PREC_WS2_TF32_F16C = 6, 7, 8, 9

ABI_VERSION = 4   # DMVS_ABI_VERSION of include/diffmvs_b200.h

f32p = C.c_void_p
i32 = C.c_int32


class ConvDesc(C.Structure):
    """Mirror of `dmvs_conv_desc` (include/diffmvs_b200.h)."""
    _fields_ = [
        ("x", f32p), ("x2", f32p),
        ("N", i32), ("D", i32), ("H", i32), ("W", i32),
        ("C1", i32), ("C2", i32), ("x_ps", i32), ("x2_ps", i32), ("in_up2", i32),
        ("in_stats", C.c_void_p), ("in_g1", f32p), ("in_g0", f32p), ("in_inv_count", C.c_float),
        ("w", f32p), ("w_t", f32p), ("w_tc", f32p), ("w_ws", f32p), ("w_ws_pair", f32p), ("w_ws16", f32p), ("precision", i32), ("bias", f32p),
        ("KD", i32), ("KH", i32), ("KW", i32), ("stride", i32), ("pad_d", i32), ("pad_h", i32), ("pad_w", i32),
        ("y", f32p), ("Do", i32), ("Ho", i32), ("Wo", i32), ("Cout", i32), ("y_ps", i32),
        ("act", i32), ("act_c0", i32), ("res_mode", i32), ("res", f32p), ("res_ps", i32), ("res_up2", i32),
        ("epi", i32), ("aux1", f32p), ("aux2", f32p), ("aux1_ps", i32), ("aux2_ps", i32), ("gru_hidden", i32),
        ("explicit_extent", i32), ("y_row_stride", i32), ("res_row_stride", i32),
        ("out_stats", C.c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/diffmvs_b200.h
SIGNATURES = {
    "dmvs_abi_version": (C.c_int, []),
    "dmvs_build_info": (C.c_char_p, []),
    "dmvs_launch_count": (C.c_uint64, []),
    "dmvs_conv_f32": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "dmvs_conv_backends": (C.c_int, [C.POINTER(ConvDesc)]),
    "dmvs_conv_ws_plan": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), i32]),
    "dmvs_conv_ws2_plan": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), i32]),
    "dmvs_conv_ws2_timeline": (C.c_int, [C.POINTER(C.c_int64), i32]),
    "dmvs_deconv3d_f32": (C.c_int, [f32p, f32p, f32p, f32p, f32p, i32, i32, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_conv3d_to1_f32": (C.c_int, [f32p, i32, C.c_void_p, C.c_float, f32p, i32, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_border_bias_add": (C.c_int, [f32p, i32, f32p, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_compose_homographies": (C.c_int, [f32p, f32p, i32, i32, C.c_void_p]),
    "dmvs_warp_volume": (C.c_int, [f32p, i32, f32p, f32p, f32p, i32, i32, i32, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_plane_sweep_corr": (C.c_int, [f32p, f32p, f32p, f32p, i32, i32, i32, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_view_weight_max": (C.c_int, [f32p, f32p, i32, i32, i32, C.c_void_p]),
    "dmvs_aggregate_views": (C.c_int, [f32p, f32p, f32p, i32, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_depth_regression": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, C.c_void_p, i32, i32, i32, C.c_void_p]),
    "dmvs_get_cost": (C.c_int, [f32p, f32p, f32p, f32p, i32, f32p, f32p, f32p, f32p, i32, f32p, i32, i32, i32, i32, i32,
                                i32, i32, i32, i32, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "dmvs_groupnorm_silu_add": (C.c_int, [f32p, C.c_void_p, f32p, f32p, f32p, i32, f32p, i32, i32, i32, i32,
                                          C.c_void_p]),
    "dmvs_upsample_depth": (C.c_int, [f32p, f32p, i32, f32p, f32p, f32p, f32p, f32p, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_refine_update": (C.c_int, [i32, f32p, f32p, i32, C.c_float, f32p, f32p, f32p, i32, f32p, f32p, f32p, i32,
                                     i32, C.c_void_p]),
    "dmvs_ddim_step": (C.c_int, [f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int64, C.c_void_p]),
    "dmvs_upsample_nearest": (C.c_int, [f32p, i32, f32p, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_image_to_nhwc4": (C.c_int, [f32p, f32p, i32, i32, C.c_void_p]),
    "dmvs_image_u8_to_nhwc4": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, i32, f32p, i32, i32, C.c_void_p]),
    "dmvs_geo_consistency": (C.c_int, [f32p, f32p, C.c_void_p, C.c_float, C.c_float, C.c_double, C.c_float, C.c_void_p, f32p,
                                       f32p, f32p, f32p, C.c_void_p, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_fuse_points": (C.c_int, [f32p, f32p, C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, f32p, i32, i32, C.c_void_p]),
    "dmvs_fuse_view": (C.c_int, [f32p, C.c_void_p, C.c_void_p, i32, i32, i32, i32, i32, C.c_void_p, C.c_void_p, i32,
                                 C.c_void_p, C.c_float, C.c_float, C.c_double, C.c_float, i32, i32, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, f32p, C.c_void_p]),
    "dmvs_nchw_to_nhwc": (C.c_int, [f32p, f32p, i32, i32, i32, i32, C.c_void_p]),
    "dmvs_nhwc_to_nchw": (C.c_int, [f32p, i32, f32p, i32, i32, i32, C.c_void_p]),
}

_LIB: Optional[C.CDLL] = None


class KernelLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the kernel library; raises if it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise KernelLibraryError(
                f"{LIB_PATH} is missing - run `python -m diffmvs_b200.build`; there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.dmvs_abi_version() != ABI_VERSION:
            raise KernelLibraryError("ABI version mismatch between _cabi.py and the built library")
        _LIB = handle
    return _LIB


_ERR = {-1: "invalid argument", -2: "misaligned pointer or stride", -3: "unsupported configuration"}


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise ValueError(f"{what}: {_ERR.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")
