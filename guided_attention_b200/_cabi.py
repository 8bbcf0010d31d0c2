"""ctypes binding of `include/guided_attn.h` (libguidedattn.so).

The library is built in-tree by `__graft_entry__.build()` (or `python -m guided_attention_b200.build`).  Loading is
lazy and failure is loud: there is no CPU / PyTorch fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libguidedattn.so")

GA_OK = 0
GA_F32, GA_F16, GA_BF16 = 0, 1, 2
GA_IMPL_AUTO, GA_IMPL_SIMT, GA_IMPL_TCGEN05, GA_IMPL_TCGEN05_SINGLE, GA_IMPL_TCGEN05_PIPE = 0, 1, 2, 3, 4
GA_TOKEN_COOR, GA_TOKEN_BOX, GA_TOKEN_KEYWORD = 0, 1, 2
GA_MAX_ACC_SLICES, GA_MAX_TOKENS, GA_MAX_BOXES, GA_MAX_CTX = 32, 24, 32, 128
(GA_STAT_MAX, GA_STAT_SUM, GA_STAT_COL, GA_STAT_ROW, GA_STAT_INSIDE, GA_STAT_OUTSIDE, GA_STAT_SCALED,
 GA_STAT_UNSCALED, GA_STAT_HINGE_IN, GA_STAT_HINGE_OUT, GA_STAT_NINSIDE, GA_STAT_CENTER, GA_STAT_RAW_SUM,
 GA_STAT_RAW_COL, GA_STAT_RAW_ROW) = range(15)
GA_STATS = 16
GA_ABI_VERSION = 7
GA_STEP_CTL_BYTES, GA_STEP_COUNTER_BASE = 256, 19
(GA_STEP_N_EVAL, GA_STEP_N_UPDATE, GA_STEP_N_CFG, GA_STEP_N_REFINE, GA_STEP_N_ROUNDS, GA_STEP_N_RENOISE) = range(6)


class GaToken(C.Structure):
    _fields_ = [("column", C.c_int32), ("kind", C.c_int32), ("box", C.c_int32), ("group", C.c_int32),
                ("target_x", C.c_float), ("target_y", C.c_float), ("center_weight", C.c_float),
                ("group_weight", C.c_float)]


class GaTailParams(C.Structure):
    _fields_ = [("res", C.c_int32), ("n_ctx", C.c_int32), ("first", C.c_int32), ("last", C.c_int32),
                ("n_tokens", C.c_int32), ("n_groups", C.c_int32), ("strict", C.c_int32), ("smooth", C.c_int32),
                ("w1d", C.c_float * 3), ("temperature", C.c_float), ("inv_count", C.c_float),
                ("inside_scale", C.c_float), ("outside_scale", C.c_float), ("n_samples", C.c_int32)]


class GaScoreBias(C.Structure):
    _fields_ = [("mask", C.c_void_p), ("mask_stride_bh", C.c_int64), ("mask_stride_n", C.c_int64),
                ("pww_masks", C.c_void_p), ("pww_coef", C.c_void_p), ("pww_smax", C.c_void_p),
                ("pww_count", C.c_int32), ("pww_column", C.c_int32 * GA_MAX_TOKENS)]


class GaStepPrograms(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("eval", "update", "cfg", "advance", "renoise", "refine_update")]


class GaStepParams(C.Structure):
    _fields_ = [("thr_call", C.c_double), ("thr_cfg", C.c_double), ("thr_last", C.c_double),
                ("has_thr_call", C.c_int32), ("has_thr_cfg", C.c_int32), ("has_thr_last", C.c_int32),
                ("check", C.c_int32), ("update_cond", C.c_int32), ("recurse_ok", C.c_int32),
                ("renoise_ok", C.c_int32), ("recurse_steps", C.c_int32), ("max_refine", C.c_int32),
                ("timestep", C.c_int64), ("step_size", C.c_float), ("ddim", C.c_float * 4),
                ("renoise", C.c_float * 2)]


_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
_bias = C.POINTER(GaScoreBias)

# name -> (restype, argtypes); every symbol declared in include/guided_attn.h
PROTOTYPES = {
    "ga_version": (_i, []),
    "ga_last_error": (C.c_char_p, []),
    "ga_device_supported": (_i, [_i]),
    "ga_cross_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "ga_cross_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i,
                                _vp]),
    "ga_attn_probs": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ga_cross_attn_smax": (_i, [_vp, _vp, _vp, _bias, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ga_cross_attn_fwd_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _bias, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "ga_cross_attn_bwd_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _bias, _vp, _i, _i, _i, _i, _i,
                                   _f, _i, _i, _vp]),
    "ga_attn_probs_ex": (_i, [_vp, _vp, _vp, _bias, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ga_self_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "ga_self_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "ga_rasterize_boxes": (_i, [C.POINTER(C.c_double), _i, _i, C.c_double, _vp, _vp]),
    "ga_guidance_tail_fwd": (_i, [C.POINTER(_vp), C.POINTER(C.c_int32), _i, C.POINTER(GaTailParams), C.POINTER(GaToken),
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ga_guidance_tail_bwd": (_i, [C.POINTER(GaTailParams), C.POINTER(GaToken), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _i, _vp]),
    "ga_step_driver_create": (_i, [C.POINTER(_vp), C.POINTER(GaStepPrograms), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   C.POINTER(GaToken), _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ga_step_driver_run": (_i, [_vp, C.POINTER(GaStepParams), _vp]),
    "ga_step_driver_set_params": (_i, [_vp, C.POINTER(GaStepParams), _vp]),
    "ga_step_driver_destroy": (_i, [_vp]),
    "ga_smooth_fwd": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_float), _vp]),
    "ga_smooth_bwd": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_float), _vp]),
    "ga_box_loss_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ga_box_loss_bwd": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ga_group_norm_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "ga_group_norm_fwd": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "ga_group_norm_bwd": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ga_add_bias_residual": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "ga_layer_norm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _i, _vp]),
    "ga_geglu_fwd": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "ga_geglu_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
}

_lib = None
_lock = threading.Lock()


class GuidedAttnLibraryError(RuntimeError):
    pass


def load(path: str = None) -> C.CDLL:
    """dlopen libguidedattn.so and attach prototypes.  Raises if the library is missing: no fallback exists."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("GA_LIB_PATH", LIB_PATH)
        if not os.path.isfile(p):
            raise GuidedAttnLibraryError(
                f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The guidance path has no CPU or PyTorch fallback.")
        lib = C.CDLL(p)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)   # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.ga_version() != GA_ABI_VERSION:
            raise GuidedAttnLibraryError(f"ABI mismatch: library {lib.ga_version()} vs binding {GA_ABI_VERSION}")
        if path is None:
            _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != GA_OK:
        msg = load().ga_last_error().decode("utf-8", "replace")
        raise GuidedAttnLibraryError(f"{what} failed ({rc}): {msg}")
