"""ctypes binding of libsug_b200.so (C ABI: include/sug_b200.h).

There is no CPU fallback: if the library is missing or cannot be built, importing the ops fails
loudly.  The library is built in-tree by ``sug_b200.build`` (nvcc, sm_100a) so it travels with the
repository snapshot and shows up as a loaded ``.so`` of this process.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsug_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sug_b200.h")

P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

# name -> (restype, argtypes); must list every function declared in include/sug_b200.h
SIGNATURES = {
    "sug_version": (I, []),
    "sug_last_error": (c_char_p, []),
    "sug_knn_ws_bytes": (Z, [I, I, I, I]),
    "sug_knn_f32": (I, [P, I, I, I, I, L, L, L, P, P, Z, P]),
    "sug_knn_reverse": (I, [P, I, I, I, P, P, P]),
    "sug_edgeconv_ws_bytes": (Z, [I, I, I, I, I]),
    "sug_edgeconv_fwd": (I, [P, L, P, P, P, P, P, P, I, I, I, I, I, F, F, F, I, P, L, P, P, P, P, P, P, Z, P]),
    "sug_edgeconv_bwd": (I, [P, L, P, L, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, P, L, I, P, P, P, P, P,
                             Z, P]),
    "sug_mlp_pool_ws_bytes": (Z, [I, I, I, I]),
    "sug_mlp_pool_fwd": (I, [P, L, P, P, P, P, P, P, I, I, I, I, F, F, F, I, I, P, P, P, P, P, Z, P]),
    "sug_mlp_pool_bwd": (I, [P, P, L, P, P, P, P, P, P, P, I, I, I, I, F, I, P, L, I, P, P, P, P, P, Z, P]),
    "sug_mmd_ws_bytes": (Z, [I, I]),
    "sug_mmd_rbf_fwd": (I, [P, L, I, I, P, I, P, I, P, P, P, Z, P]),
    "sug_mmd_rbf_bwd": (I, [P, L, I, I, P, P, P, L, P]),
    "sug_chamfer_f32": (I, [P, P, I, I, I, P, P, P]),
    "sug_fps": (I, [P, I, I, I, P, P, P]),
    "sug_ball_query": (I, [P, P, I, I, I, F, I, P, P]),
    "sug_knn_query": (I, [P, P, I, I, I, I, P, P]),
    "sug_three_nn": (I, [P, P, I, I, I, I, P, P]),
    "sug_knn_query_set": (I, [P, P, I, I, I, I, P, P]),
    "sug_group_max_fwd": (I, [P, P, I, I, I, I, I, P, P, P]),
    "sug_group_max_bwd": (I, [P, P, I, I, I, I, P, P]),
    "sug_interp_fwd": (I, [P, P, P, I, I, I, I, I, P, P]),
    "sug_interp_bwd": (I, [P, P, P, P, I, I, I, I, I, P, P, P]),
    "sug_gemm_f32": (I, [P, L, L, P, L, L, P, P, L, I, I, I, I, P]),
    "sug_gemm_auto_f32": (I, [P, L, L, P, L, L, P, P, L, I, I, I, P]),
    "sug_linear_bn_act_fwd": (I, [P, L, P, P, P, P, P, P, L, I, I, F, F, F, I, P, P, L, P, P, Z, P]),
    "sug_linear_bn_act_bwd": (I, [P, L, P, L, P, P, P, P, P, L, I, I, F, P, L, P, P, P, P, P, Z, P]),
    "sug_gemm_tc_f32": (I, [P, L, I, P, L, I, P, P, L, I, I, I, P]),
    "sug_node_offset_fwd": (I, [P, P, P, P, I, I, I, I, P, P]),
    "sug_node_offset_bwd": (I, [P, P, P, P, P, I, I, I, I, P, P]),
    "sug_interp_weight_fwd": (I, [P, P, P, I, I, I, I, P, P]),
    "sug_interp_weight_bwd": (I, [P, P, P, P, I, I, I, I, P, P]),
    "sug_sda_sem_weights": (I, [P, P, P, P, I, I, F, P, P]),
    "sug_soft_mmd_assemble": (I, [P, L, P, L, P, P, I, I, I, F, P, P]),
    "sug_focal_loss_fwd": (I, [P, P, P, I, I, F, I, P, P]),
    "sug_focal_loss_bwd": (I, [P, P, P, P, I, I, F, I, P, P]),
    "sug_adam_chunk": (I, []),
    "sug_adam_f32": (I, [P, P, P, P, P, P, P, I, L, P, P, F, F, F, F, P]),
    "sug_adam_multi_f32": (I, [P, P, P, P, P, P, P, P, P, P, I, L, P, I, P]),
    "sug_prof_num_classes": (I, []),
    "sug_prof_class_name": (c_char_p, [I]),
    "sug_prof_enable": (None, [ctypes.c_uint]),
    "sug_prof_reset": (None, []),
    "sug_prof_collect": (I, [P, P, P, P, P]),
}


def header_symbols():
    """Names of all functions declared in include/sug_b200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sug_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build
        build.build_library()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m sug_b200.build` (needs nvcc). "
                           "sug_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SugError(RuntimeError):
    pass


def check(status: int, what: str):
    if status != 0:
        msg = load().sug_last_error().decode(errors="replace")
        raise SugError(f"{what} failed with status {status}: {msg}")


def prof_reset(mask: int = 0):
    lib = load()
    lib.sug_prof_reset()
    lib.sug_prof_enable(mask)


def prof_collect():
    """{class name: dict(ms, timed, launches, flops, bytes)} since the last prof_reset()."""
    lib = load()
    n = lib.sug_prof_num_classes()
    D, LL = ctypes.c_double * n, ctypes.c_longlong * n
    ms, timed, launches, flops, byts = D(), LL(), LL(), D(), D()
    check(lib.sug_prof_collect(ms, timed, launches, flops, byts), "sug_prof_collect")
    out = {}
    for i in range(n):
        out[lib.sug_prof_class_name(i).decode()] = dict(index=i, ms=ms[i], timed=timed[i], launches=launches[i],
                                                        flops=flops[i], bytes=byts[i])
    return out
