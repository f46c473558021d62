"""ctypes binding of libv2x_b200.so (the C ABI declared in include/v2x_b200.h).

The library is the product: there is no python / torch / CPU fallback.  Importing this module
without the built library, or calling an op without a Blackwell GPU, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libv2x_b200.so")

EPI_ACT, EPI_F32_SPLIT, EPI_GRU, EPI_F32_NCHW, EPI_TAIL_F32_SPLIT = 0, 1, 2, 3, 4


class ConvParams(C.Structure):
    """Mirror of ``v2x_conv_params`` (include/v2x_b200.h)."""
    _fields_ = [
        ("src", C.c_void_p * 2),
        ("cin", C.c_int32 * 2),
        ("n_maps", C.c_int32),
        ("h_out", C.c_int32),
        ("w_out", C.c_int32),
        ("stride", C.c_int32),
        ("taps", C.c_int32),
        ("planes", C.c_int32),
        ("weights", C.c_void_p),
        ("bias", C.c_void_p),
        ("cout", C.c_int32),
        ("cout_pad", C.c_int32),
        ("block_n", C.c_int32),
        ("epilogue", C.c_int32),
        ("relu", C.c_int32),
        ("upsample2x", C.c_int32),
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("out_c_total", C.c_int32),
        ("out_c_off", C.c_int32),
        ("split", C.c_int32),
        ("gru_bhn", C.c_void_p),
        ("passthrough", C.c_void_p),
        ("num_agent", C.c_void_p),
        ("batch", C.c_int32),
        ("agents", C.c_int32),
        ("map_offset", C.c_int32),
        ("gru_pre_act", C.c_int32),
        ("tap_pack", C.c_int32),
        ("mmas", C.c_int32),
        ("gru_add", C.c_void_p),
        ("tail_weights", C.c_void_p),
        ("tail_bias", C.c_void_p),
        ("tail_cout", C.c_int32),
        ("tail_cout_pad", C.c_int32),
    ]


# every symbol include/v2x_b200.h declares: (name, restype, argtypes)
_I32, _I64, _F32, _P = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = [
    ("v2x_version", C.c_int, []),
    ("v2x_last_error", C.c_char_p, []),
    ("v2x_device_ok", C.c_int, []),
    ("v2x_conv_fwd", C.c_int, [C.POINTER(ConvParams), _P]),
    ("v2x_conv_fwd_crosscheck", C.c_int, [C.POINTER(ConvParams), _P]),
    ("v2x_set_debug_mode", C.c_int, [C.c_int]),
    ("v2x_pack_conv_weights", C.c_int,
     [_P, _P, _P, _P, _P, _P, _F32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _I32, _I32,
      _I32, _I32, _P]),
    ("v2x_pack_gru_bias", C.c_int, [_P, _P, _I32, _P, _P, _P]),
    ("v2x_pack_input", C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _P]),
    ("v2x_warp_mean_fwd", C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_act_to_nchw_f32", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_linear_fwd", C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_pack_input_nchw", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_maxpool2_fwd", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_upsample_bilinear2_fwd", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_attn_scores_fwd", C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_warp_gated_fwd", C.c_int, [_P, _P, _P, _P, _P] + [_I32] * 12 + [_P]),
    ("v2x_warp_reduce_fwd", C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_pair_score_fwd", C.c_int, [_P] * 10 + [_I32] * 6 + [_P]),
    ("v2x_agent_softmax_fwd", C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    ("v2x_warp_weighted_fwd", C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_voxelize_fwd", C.c_int, [_P, _P, _I32, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    ("v2x_pack_input_u8", C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _P]),
    ("v2x_det_nms_fwd", C.c_int, [_P, _P, _P, _I32, _I32, _I32, _F32, _F32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("v2x_restore_absent_fwd", C.c_int, [_P, _P, _P, _I32, _I32, _I64, _I32, _P]),
    # training step
    ("v2x_bn_stats_fwd", C.c_int, [_P, _I64, _I32, _I32, _P, _P, _P]),
    ("v2x_bn_finalize", C.c_int, [_P, _P, _I64, _P, _P, _F32, _F32, _P, _P, _P, _P, _P, _P, _I32, _P]),
    ("v2x_bn_relu_apply_fwd", C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _I32, _P]),
    ("v2x_bn_relu_bwd", C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _I32, _P, _P, _P]),
    ("v2x_resample2", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_act_add", C.c_int, [_P, _P, _I64, _I32, _P]),
    ("v2x_conv_wgrad", C.c_int, [_P, _P] + [_I32] * 8 + [_P] + [_I32] * 4 + [_F32, _P]),
    ("v2x_conv_wgrad_tc", C.c_int, [_P, _P] + [_I32] * 8 + [_P] + [_I32] * 4 + [_F32, _P]),
    ("v2x_scale_to_f32", C.c_int, [_P, _P, _I32, _F32, _I32, _P]),
    ("v2x_gru_gates_fwd", C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _I32, _I32, _P]),
    ("v2x_gru_gates_bwd", C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _I32, _I32, _P, _P]),
    ("v2x_warp_mean_bwd", C.c_int, [_P, _P, _P, _P] + [_I32] * 8 + [_P]),
    ("v2x_warp_reduce_bwd", C.c_int, [_P, _P, _P, _P, _P, _P] + [_I32] * 8 + [_P]),
    ("v2x_warp_weighted_bwd", C.c_int, [_P, _P, _P, _P, _P, _P, _P] + [_I32] * 7 + [_P]),
    ("v2x_warp_gated_bwd", C.c_int, [_P, _P, _P, _P, _P, _P, _P] + [_I32] * 8 + [_P]),
    ("v2x_maxpool2_bwd", C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    ("v2x_upsample_bilinear2_bwd", C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    # x_3 exchange over NVLink peer memory
    ("v2x_peer_alloc", C.c_int, [_I64, C.POINTER(_P), _P]),
    ("v2x_peer_open", C.c_int, [_P, C.POINTER(_P)]),
    ("v2x_peer_close", C.c_int, [_P]),
    ("v2x_peer_free", C.c_int, [_P]),
    ("v2x_peer_begin", C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    ("v2x_peer_push", C.c_int, [_P, _I64, _I64, C.POINTER(_P), C.POINTER(_I32), _I64, C.POINTER(_P), _P, _P, _I32, _I32, _P]),
    ("v2x_peer_wait", C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    ("v2x_peer_done", C.c_int, [C.POINTER(_P), _P, _I32, _I32, _P]),
]

_lib = None


class V2XError(RuntimeError):
    pass


def load():
    """Load libv2x_b200.so (built in-tree by ``__graft_entry__.build()``); raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise V2XError(
            "libv2x_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU / torch fallback for this path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().v2x_last_error().decode("utf-8", "replace")
        raise V2XError("%s failed (%d): %s" % (what or "v2x call", rc, msg))
