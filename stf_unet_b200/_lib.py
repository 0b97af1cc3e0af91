"""ctypes binding of libstfb200.so (the C ABI declared in include/stfb200.h).

There is no fallback: if the shared library has not been built (``python -m stf_unet_b200.build`` or
``__graft_entry__.build()``) importing an op raises, and every launch on a machine without an sm_100
device returns STFB_ENODEV which is raised as RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstfb200.so")

F32, BF16, BF16X3 = 0, 1, 2
CONV_FWD, CONV_TRANSPOSED = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2


class ConvParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x", "x2", "w", "y", "bias", "bias2", "scale", "shift", "residual")] + \
               [(n, C.c_int) for n in ("N", "H", "W", "C1", "C2", "Ho", "Wo", "Cout", "kh", "kw", "stride", "pad",
                                       "ldw", "mode", "relu", "x_dtype", "y_dtype", "impl")] + \
               [("stat_partial", C.c_void_p), ("stat_slots", C.c_int), ("stat_groups", C.c_int)]


_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_SIZE_T_FUNCS = {"stfb_conv2d_wgrad_workspace_bytes": [_vp, _vp] + [_i] * 14, "stfb_wgrad_scratch_bytes": [],
                 "stfb_bn_bwd_fused_scratch_floats": [_i, _i]}
_SIGS = {
    "stfb_conv2d": [C.POINTER(ConvParams), _vp],
    "stfb_conv2d_tcgen05_supported": [C.POINTER(ConvParams)],
    "stfb_conv2d_stats_fusable": [C.POINTER(ConvParams), _i],
    "stfb_conv2d_wgrad": [_vp, _vp, _vp] + [_i] * 15 + [_vp, C.c_size_t, _vp],
    "stfb_conv2d_wgrad_tcgen05_supported": [_vp, _vp] + [_i] * 12,
    "stfb_pack_weight": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "stfb_pack_weight_ex": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "stfb_split_bf16x3": [_vp, _vp, _ll, _i, _vp],
    "stfb_pack_weight_split": [_vp, _vp] + [_i] * 6 + [_vp],
    "stfb_lstm_step_fused": [_vp] * 9 + [_i, _i, _i, _i, _vp],
    "stfb_set_wgrad_scratch": [_vp, C.c_size_t],
    "stfb_wgrad_scatter_batched": [_vp, _i, _ll, _vp, _vp, _vp],
    "stfb_pack_weights_batched": [_vp, _i, _ll, _i, _vp],
    "stfb_im2col_small": [_vp, _vp] + [_i] * 10 + [_vp],
    "stfb_unpad_wgrad": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "stfb_bn_partial_blocks": [_i, _ll],
    "stfb_bn_stats": [_vp, _vp, _i, _i, _ll, _i, _i, _vp],
    "stfb_bn_finalize_train": [_vp, _i] + [_vp] * 9 + [_i, _ll, _i, _f, _f, _vp],
    "stfb_bn_fold_eval": [_vp] * 6 + [_i, _f, _vp],
    "stfb_bn_apply": [_vp] * 5 + [_i, _ll, _i, _i, _i, _vp],
    "stfb_bn_relu_maxpool_from_stats": [_vp, _vp, _i, _vp, _vp, _vp, _vp] + [_i] * 10 + [_f, _i, _vp],
    "stfb_bn_apply_from_stats": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _ll, _i, _f, _i, _i, _vp],
    "stfb_bn_bwd_reduce": [_vp] * 8 + [_i, _i, _ll, _i, _i, _i, _vp],
    "stfb_bn_bwd_finalize": [_vp, _i] + [_vp] * 5 + [_i, _ll, _i, _vp],
    "stfb_bn_bwd_apply": [_vp] * 9 + [_i, _i, _ll, _i, _i, _i, _vp],
    "stfb_bn_bwd_fused": [_vp] * 12 + [_i, _i, _ll, _i, _i, _i, _vp],
    "stfb_colsum": [_vp, _vp, _ll, _i, _i, _vp],
    "stfb_maxpool_fwd": [_vp, _vp] + [_i] * 10 + [_vp],
    "stfb_maxpool_bwd": [_vp, _vp, _vp] + [_i] * 10 + [_vp],
    "stfb_maxpool_fwd_idx": [_vp, _vp, _vp] + [_i] * 10 + [_vp],
    "stfb_maxpool_bwd_idx": [_vp, _vp, _vp] + [_i] * 10 + [_vp],
    "stfb_bilinear_fwd": [_vp, _vp] + [_i] * 7 + [_vp],
    "stfb_bilinear_bwd": [_vp, _vp] + [_i] * 7 + [_vp],
    "stfb_lstm_cell_fwd": [_vp] * 5 + [_ll, _i, _i, _vp],
    "stfb_lstm_cell_bwd": [_vp] * 6 + [_ll, _i, _i, _i, _vp],
    "stfb_pack_series": [_vp, _vp] + [_i] * 6 + [_vp],
    "stfb_pack_series_maps": [_vp, _vp, _vp] + [_i] * 7 + [_vp],
    "stfb_pack_series_u8": [_vp, _vp] + [_i] * 4 + [_f, _f, _i, _vp],
    "stfb_adamw_flat": [_vp] * 4 + [_ll] + [C.c_double] * 5 + [_ll, C.c_double, _vp],
    "stfb_repeat": [_vp, _vp, C.c_size_t, _i, _vp],
    "stfb_nhwc_to_nchw": [_vp, _vp] + [_i] * 5 + [_vp],
    "stfb_nchw_to_nhwc": [_vp, _vp] + [_i] * 5 + [_vp],
    "stfb_add_inplace": [_vp, _vp, _ll, _i, _vp],
    "stfb_cast": [_vp, _i, _vp, _i, _ll, _vp],
    "stfb_ce_dice_fwd": [_vp] * 4 + [_i, _i, _i, _f, _vp],
    "stfb_eval_metrics": [_vp] * 7 + [_i, _i, _i, _ll, _i, _vp],
    "stfb_ce_dice_bwd": [_vp] * 5 + [_i, _i, _i, _f, _vp],
    "stfb_ce_dice_fwd_ex": [_vp] * 5 + [_i, _i, _i, _f, _ll, _i, _vp],
    "stfb_ce_dice_bwd_ex": [_vp] * 6 + [_i, _i, _i, _f, _ll, _i, _vp],
    "stfb_lstm_bwd_step_fused": [_vp] * 7 + [_i] * 4 + [_vp],
    "stfb_lstm_seq_supported": [_i] * 5,
    "stfb_lstm_seq_fused": [_vp] * 7 + [_i] * 6 + [_vp],
    "stfb_augment_series_u8": [_vp] * 6 + [_i] * 6 + [_f, _f, _vp],
    "stfb_tofts_forward": [_vp] * 5 + [_i, _i, _f] + [_vp] * 4 + [_ll, _vp],
    "stfb_tofts_fit": [_vp] * 6 + [_i, _i, _f] + [_vp] * 3 + [_ll, _i, _i, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp],
}
EXPORTS = sorted(list(_SIGS) + list(_SIZE_T_FUNCS) + ["stfb_version", "stfb_last_error", "stfb_launch_count"])

_lib = None


def load():
    """Load libstfb200.so; raises RuntimeError (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -m stf_unet_b200.build` (or __graft_entry__.build()); "
                           "stf_unet_b200 has no CPU / PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = C.c_int
    for name, sig in _SIZE_T_FUNCS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = C.c_size_t
    lib.stfb_version.restype = C.c_int
    lib.stfb_last_error.restype = C.c_char_p
    lib.stfb_launch_count.restype = C.c_ulonglong
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().stfb_last_error().decode(errors="replace")
        raise RuntimeError(f"libstfb200 {what} failed ({status}): {msg}")


def launch_count():
    return int(load().stfb_launch_count())
