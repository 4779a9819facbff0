"""ctypes binding of libstcgan_b200.so (the C ABI declared in include/stcgan_b200.h).

The library is the ONLY arithmetic provider of this package: there is no eager/PyTorch/CPU
fallback.  If the shared object is missing or fails to load, importing any compute entry
point raises `StcganLibraryError` loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstcgan_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
GEOM_WIN_S2, GEOM_WIN_S1, GEOM_WIN_S1_FLIP, GEOM_PARITY = 0, 1, 2, 3
BACKEND_FFMA, BACKEND_TC = 0, 1


class StcganLibraryError(RuntimeError):
    pass


class StcganError(RuntimeError):
    pass


class LossTerm(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("grad", C.c_void_p), ("n", C.c_int64),
                ("target", C.c_float), ("weight", C.c_float), ("kind", C.c_int32), ("slot", C.c_int32),
                ("accumulate", C.c_int32), ("loss_weight", C.c_float)]


class AdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n", C.c_int64), ("d0", C.c_int32), ("d1", C.c_int32), ("p1", C.c_void_p), ("p2", C.c_void_p)]


class AugSample(C.Structure):
    _fields_ = [("scale_inv", C.c_double * 6), ("rot_inv", C.c_double * 6), ("flip", C.c_int32), ("row_off", C.c_int32),
                ("col_off", C.c_int32), ("identity", C.c_int32)]


_i, _p, _f, _i64 = C.c_int, C.c_void_p, C.c_float, C.c_int64

# name -> (restype, argtypes); every symbol declared in include/stcgan_b200.h is listed here
SIGNATURES = {
    "stcgan_abi_version": (_i, []),
    "stcgan_arch": (C.c_char_p, []),
    "stcgan_error_string": (C.c_char_p, [_i]),
    "stcgan_launch_count": (_i64, []),
    "stcgan_launch_count_reset": (None, []),
    "stcgan_tapconv": (_i, [_i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _p, _i64, _p]),
    "stcgan_tapconv_bnstats": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p, _i64, _p, _p]),
    "stcgan_tapconv_ep": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i64, _p]),
    "stcgan_thinconv2": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _p]),
    "stcgan_bn_fused_apply": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _f, _f, _i, _p, _p, _i, _i,
                                   _p, _i, _i, _p, _i, _i, _p]),
    "stcgan_tapwgrad": (_i, [_i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p]),
    "stcgan_pack_weight": (_i, [_i, _p, _i, _i, _p, _p, _p]),
    "stcgan_unpack_grad": (_i, [_p, _i, _i, _p, _i, _p]),
    "stcgan_bn_stats": (_i, [_i, _p, _i64, _i, _i, _p, _p]),
    "stcgan_bn_finalize": (_i, [_p, _i64, _i, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p]),
    "stcgan_bn_running_update": (_i, [_p, _i64, _i, _p, _p, _f, _p]),
    "stcgan_bn_act_apply": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _p, _i, _i, _p]),
    "stcgan_bn_act_bwd_small": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p, _i, _i, _p, _i, _i, _p, _i, _p, _p, _p]),
    "stcgan_bn_act_bwd_reduce": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _i, _i, _p, _i, _i, _p, _p]),
    "stcgan_bn_act_bwd_apply": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _p, _i, _i, _p, _i, _i,
                                     _p, _p, _i, _p, _p, _p, _p]),
    "stcgan_colsum": (_i, [_i, _p, _i64, _i, _i, _p, _p]),
    "stcgan_pack_input": (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "stcgan_tapconv_thin_n": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p]),
    "stcgan_thinconv": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "stcgan_thinwgrad": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "stcgan_thin_col2im": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p, _p, _i, _i, _i, _p]),
    "stcgan_thin_convt_u8": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p, _p, _i, _i, _p]),
    "stcgan_pack_weight_tapn": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "stcgan_pack_weight_thin": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "stcgan_pack_weight_pad16": (_i, [_p, _i, _i, _i, _p, _p]),
    "stcgan_unpack_input_grad": (_i, [_i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "stcgan_nhwc_to_nchw": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p]),
    "stcgan_nchw_to_nhwc": (_i, [_i, _p, _i, _i, _i, _i, _p, _i, _p]),
    "stcgan_out_act_bwd": (_i, [_i, _i, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "stcgan_fused_loss": (_i, [C.POINTER(LossTerm), _i, _p, _p]),
    "stcgan_rel_logits": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p]),
    "stcgan_adam_step": (_i, [_p, _p, _i, _p, _p]),
    "stcgan_adam_step_range": (_i, [_p, _p, _i, _i, _p, _i, _i, _p]),
    "stcgan_adam_chunk": (_i, []),
    "stcgan_adam_tile": (_i, []),
    "stcgan_float2uint_hwc": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "stcgan_float2uint": (_i, [_p, _i64, _p, _p]),
    "stcgan_u8_hwc_to_nchw_f32": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "stcgan_augment_u8": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises StcganLibraryError if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StcganLibraryError(
            f"{LIB_PATH} not found: build it with `python shadow-removal-istd_b200/build.py` "
            "(or __graft_entry__.build()).  stcgan_b200 has no non-CUDA fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise StcganLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise StcganLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.stcgan_abi_version() != 1:
        raise StcganLibraryError("ABI version mismatch")
    _lib = lib
    return lib


BN_SLOTS = 4     # STCGAN_BN_SLOTS of include/stcgan_b200.h


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().stcgan_error_string(code).decode()
        raise StcganError(f"{what}: {msg} (code {code})")


def launch_count() -> int:
    return int(load().stcgan_launch_count())


def launch_count_reset():
    load().stcgan_launch_count_reset()
