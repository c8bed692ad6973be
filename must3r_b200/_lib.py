"""ctypes binding of libm3r_b200.so (the C ABI declared in include/must3r_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present, the ops raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3r_b200.so")
_lib = None


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("W", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("is_bf16", C.c_int32),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("rowbias", C.c_void_p), ("rb_period", C.c_int32), ("rb_first", C.c_int32),
        ("rope_tab", C.c_void_p), ("rope_cols", C.c_int32), ("rope_period", C.c_int32),
        ("out", C.c_void_p), ("ldc", C.c_int64),
        ("out_dtype", C.c_int32),
        ("rows_per_batch", C.c_int32),
        ("batch_stride_rows", C.c_int64),
        ("n_peer_out", C.c_int32),
        ("peer_out", C.c_void_p * 8),
        ("w_static", C.c_int32),
        ("norm_out", C.c_void_p), ("ldn", C.c_int64), ("norm_eps", C.c_float),
    ]


class GemmGroup(C.Structure):
    _fields_ = [("groups", C.c_int32), ("w_group_rows", C.c_int64), ("bias_group", C.c_int64),
                ("out", C.c_void_p * 16), ("peer_out", C.c_void_p * (16 * 8)), ("max_ctas", C.c_int32)]


class AttnArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("ldq", C.c_int64),
        ("K0", C.c_void_p), ("V0", C.c_void_p), ("ldk0", C.c_int64), ("kv_bstride0", C.c_int64), ("Nk0", C.c_int32),
        ("K1", C.c_void_p), ("V1", C.c_void_p), ("ldk1", C.c_int64), ("kv_bstride1", C.c_int64), ("Nk1", C.c_int32),
        ("O", C.c_void_p), ("ldo", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32),
        ("kv_group", C.c_int32),
        ("skip_lo", C.c_int32), ("skip_step", C.c_int32), ("skip_len", C.c_int32),
        ("is_bf16", C.c_int32),
        ("scale", C.c_float),
        ("export_o", C.c_void_p), ("export_ml", C.c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/must3r_b200.h declares
SIGNATURES = {
    "m3r_last_error": (C.c_char_p, []),
    "m3r_abi_version": (C.c_int, []),
    "m3r_debug_trace": (C.c_int, [C.c_void_p]),
    "m3r_launch_count": (C.c_longlong, []),
    "m3r_prof_enable": (None, [C.c_int]),
    "m3r_prof_read": (C.c_int, [C.POINTER(C.c_double)]),
    "m3r_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "m3r_gemm_grouped": (C.c_int, [C.POINTER(GemmArgs), C.POINTER(GemmGroup), C.c_void_p]),
    "m3r_normalize16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "m3r_layernorm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "m3r_cast16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "m3r_layernorm16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int64, C.c_int32, C.c_void_p]),
    "m3r_add_cast16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                 C.c_int32, C.c_void_p]),
    "m3r_rope_table": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "m3r_rope_2d": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                              C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "m3r_attention": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "m3r_attn_merge": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                 C.c_int32, C.c_void_p]),
    "m3r_attn_state_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "m3r_im2col16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "m3r_unpatchify": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "m3r_postprocess": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "m3r_nn_min_dist": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "m3r_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "m3r_peer_free": (C.c_int, [C.c_void_p]),
    "m3r_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "m3r_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "m3r_ipc_close": (C.c_int, [C.c_void_p]),
    "m3r_peer_bcast": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_void_p]),
    "m3r_peer_signal": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_uint32, C.c_void_p]),
    "m3r_peer_wait": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
}


def lib():
    """Load the shared library once; raise (never fall back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m must3r_b200.build` "
                               "(must3r_b200 has no CPU / PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        apply_signatures()
    return _lib


def apply_signatures():
    """(Re)apply SIGNATURES to the loaded library (model/common.py registers the whole-model entry points)."""
    if _lib is None:
        return
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(_lib, name)
        fn.restype = res
        fn.argtypes = args


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().m3r_last_error()
        raise RuntimeError(f"must3r_b200 {what} failed: {msg.decode() if msg else rc}")
