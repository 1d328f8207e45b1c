"""ctypes binding of the C-ABI in ``include/tn4ml_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no CPU or PyTorch fallback: if the library is missing, or no CUDA device is
present when a compute call is made, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

TNB_MAX_PHYS = 16
TNB_MAX_OUT = 32

KIND_MPS, KIND_SMPO = 0, 1
F32, F64 = 0, 1
PREC_SIMT, PREC_TENSOR = 0, 1
EMB_TRIG, EMB_FOURIER, EMB_POLY, EMB_RBF, EMB_LINCOMP, EMB_LEGENDRE, EMB_LAGUERRE, EMB_HERMITE, EMB_ARRAYS = range(9)
REG_LOG_FROB, REG_LOG_POW_FROB, REG_LOG_RELU_FROB, REG_QUAD_FROB = range(4)
(HEAD_NLL, HEAD_SOFTMAX_XENT, HEAD_SOFTMAX_XENT_WEIGHTED, HEAD_XENT_SOFTMAX_ARGMAX, HEAD_MSE,
 HEAD_TSN, HEAD_LOGQUAD, HEAD_QUAD) = range(8)


class TnbEmbedding(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("k", C.c_int32), ("p", C.c_int32), ("degree", C.c_int32),
        ("include_bias", C.c_int32), ("n_centers", C.c_int32), ("n_in", C.c_int32), ("cross", C.c_int32),
        ("gamma", C.c_double),
        ("centers", C.c_double * TNB_MAX_PHYS),
    ]


class TnbProblem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dtype", C.c_int32), ("precision", C.c_int32), ("L", C.c_int32),
        ("chi", C.c_int32), ("d", C.c_int32), ("n_out", C.c_int32), ("out_site", C.c_int32),
        ("spacing", C.c_int32), ("head", C.c_int32), ("emb", TnbEmbedding),
        ("class_weights", C.c_double * TNB_MAX_OUT),
    ]


# TNB_LIB points at another build of the same library (measurement builds under tools/); default: the in-tree build
LIB_PATH = os.environ.get("TNB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtn4ml_b200.so")

# every symbol include/tn4ml_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "tnb_version": (C.c_char_p, []),
    "tnb_last_error": (C.c_char_p, []),
    "tnb_init": (C.c_int, [C.c_int32]),
    "tnb_ffi_available": (C.c_int32, []),
    "tnb_pack_params": (C.c_int, [_P(TnbProblem), _P(C.c_void_p), _P(C.c_int32), _P(C.c_int32), C.c_void_p, C.c_void_p]),
    "tnb_unpack_grads": (C.c_int, [_P(TnbProblem), C.c_void_p, _P(C.c_void_p), _P(C.c_int32), _P(C.c_int32), C.c_void_p]),
    "tnb_param_count": (C.c_int64, [_P(TnbProblem)]),
    "tnb_site_offset": (C.c_int64, [_P(TnbProblem), C.c_int32]),
    "tnb_site_nout": (C.c_int32, [_P(TnbProblem), C.c_int32]),
    "tnb_workspace_bytes": (C.c_int64, [_P(TnbProblem), C.c_int64, C.c_int32]),
    "tnb_forward": (C.c_int, [_P(TnbProblem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "tnb_backward": (C.c_int, [_P(TnbProblem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                               C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "tnb_embed": (C.c_int, [_P(TnbEmbedding), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnb_adam_update": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                  C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_double, C.c_void_p]),
    "tnb_adam_update_dev": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_double, C.c_void_p,
                                      C.c_void_p]),
    "tnb_gradient_clip_scratch_bytes": (C.c_int64, [_P(TnbProblem)]),
    "tnb_gradient_clip": (C.c_int, [_P(TnbProblem), C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "tnb_backward_part": (C.c_int, [_P(TnbProblem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32]),
    "tnb_pair_grad": (C.c_int, [_P(TnbProblem), C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p]),
    "tnb_part_sites": (C.c_int32, [_P(TnbProblem), C.c_int32, C.c_int32, _P(C.c_int32), _P(C.c_int32)]),
    "tnb_norm_workspace_bytes": (C.c_int64, [_P(TnbProblem)]),
    "tnb_norm": (C.c_int, [_P(TnbProblem), C.c_void_p, C.c_void_p, C.c_int32, _P(C.c_int32), _P(C.c_double), C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tnb_scale": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "tnb_set_deterministic": (C.c_int, [C.c_int32]),
    "tnb_profile": (C.c_int, [C.c_int32]),
    "tnb_profile_read": (C.c_int32, [C.c_char_p, _P(C.c_float), C.c_int32]),
    "tnb_launch_count": (C.c_int64, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the native library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root. "
                "tn4ml_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("tn4ml_b200: " + lib().tnb_last_error().decode())
