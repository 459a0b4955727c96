"""ctypes binding of libax2d.so (the C ABI in include/ax2d.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libax2d.so")

MAX_SEG = 8
ACT_CODES = {None: 0, "none": 0, "relu": 1, "leakyrelu": 2, "elu": 3, "gelu": 4, "silu": 5}
SEG_MODES = {"sum": 0, "mean": 1, "max": 2}

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class CMat(C.Structure):
    """ax2d_cmat / ax2d_mat: a column-segmented row-major fp32 matrix."""
    _fields_ = [("ptr", c_void_p * MAX_SEG), ("ld", c_int64 * MAX_SEG),
                ("width", C.c_int32 * MAX_SEG), ("n_seg", C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [("bias", c_void_p), ("pre", CMat), ("act", C.c_int32), ("act_cols", C.c_int32),
                ("mask", c_void_p), ("ld_mask", c_int64), ("drop_p", c_float), ("drop_seed", C.c_uint64),
                ("drop_tick", c_void_p),
                ("resid", CMat), ("dact_pre", c_void_p), ("ld_dact", c_int64), ("dact", C.c_int32), ("dact_cols", C.c_int32),
                ("accumulate", C.c_int32)]


_SIGNATURES = {
    "ax2d_abi_version": (c_int, []),
    "ax2d_error_string": (C.c_char_p, [c_int]),
    "ax2d_last_error": (C.c_char_p, []),
    "ax2d_launch_count": (C.c_uint64, []),
    "ax2d_set_pdl": (c_int, [c_int]),
    "ax2d_host_csr_build": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                    c_void_p, c_void_p, c_void_p]),
    "ax2d_host_csr_rows_unique": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "ax2d_host_tile_plan": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    C.POINTER(c_int64), C.POINTER(c_int64), C.POINTER(C.c_int32)]),
    "ax2d_host_shell_edges": (c_int64, [c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64]),
    "ax2d_host_shell_csr": (c_int64, [c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64]),
    "ax2d_shell_csr_count": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "ax2d_shell_csr_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "ax2d_agg": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                         c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "ax2d_agg_tiles_mma_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "ax2d_agg_tiles_mma": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "ax2d_attn_pool_fwd_config": (c_int, [c_int, c_int]),
    "ax2d_attn_pool_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ax2d_attn_pool_bwd_workspace": (c_int64, [c_int64, c_int, c_int]),
    "ax2d_attn_pool_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p]),
    "ax2d_seg_reduce_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ax2d_seg_reduce_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "ax2d_charge_eq_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_void_p]),
    "ax2d_charge_eq_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p,
                                   c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "ax2d_tetra_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int64, c_void_p]),
    "ax2d_tetra_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                               c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "ax2d_cistrans": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_void_p, c_int64, c_void_p]),
    "ax2d_embed_fwd": (c_int, [C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int64, c_void_p, c_int64,
                               c_void_p]),
    "ax2d_embed_bwd_all_workspace": (c_int64, [c_int64, c_int64, c_int]),
    "ax2d_embed_bwd_all": (c_int, [c_void_p, c_int64, c_int, c_int, c_int64, C.POINTER(c_void_p), C.POINTER(c_int64),
                                   C.POINTER(c_void_p), c_void_p, c_void_p]),
    "ax2d_embed_fwd_bf16": (c_int, [C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int64, c_void_p, c_int64,
                                    c_void_p]),
    "ax2d_embed_bwd_all_bf16": (c_int, [c_void_p, c_int64, c_int, c_int, c_int64, C.POINTER(c_void_p), C.POINTER(c_int64),
                                        C.POINTER(c_void_p), c_void_p, c_void_p]),
    "ax2d_embed_bwd_workspace": (c_int64, [c_int64, c_int]),
    "ax2d_embed_bwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "ax2d_gemm_workspace": (c_int64, [c_int64, c_int64, c_int64, c_int, c_int]),
    "ax2d_gemm": (c_int, [C.POINTER(CMat), c_int, C.POINTER(CMat), c_int, C.POINTER(CMat), c_int64, c_int64, c_int64,
                          C.POINTER(Epilogue), c_int, c_void_p, c_void_p]),
    "ax2d_split_tf32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "ax2d_gemm_tc_supported": (c_int, [C.POINTER(CMat), c_int64, c_int64, c_int64]),
    "ax2d_gemm_tc": (c_int, [C.POINTER(CMat), c_void_p, c_void_p, c_int64, C.POINTER(CMat), c_int64, c_int64, c_int64,
                             C.POINTER(Epilogue), c_void_p]),
    "ax2d_gemm_tc_wgrad_supported": (c_int, [C.POINTER(CMat), C.POINTER(CMat), c_int64, c_int64, c_int64]),
    "ax2d_gemm_tc_wgrad_splits": (c_int, [c_int64, c_int64, c_int64]),
    "ax2d_gemm_tc_wgrad_workspace": (c_int64, [c_int64, c_int64, c_int64]),
    "ax2d_gemm_tc_wgrad": (c_int, [C.POINTER(CMat), C.POINTER(CMat), C.POINTER(CMat), c_int64, c_int64, c_int64, c_int,
                                   c_void_p, c_void_p, c_void_p]),
    "ax2d_gemm_bf16_supported": (c_int, [C.POINTER(CMat), c_int64, c_int64, c_int64]),
    "ax2d_gemm_bf16": (c_int, [C.POINTER(CMat), c_void_p, c_int64, C.POINTER(CMat), c_int, c_int64, c_int64, c_int64,
                               C.POINTER(Epilogue), c_void_p]),
    "ax2d_gemm_bf16_wgrad_supported": (c_int, [C.POINTER(CMat), C.POINTER(CMat), c_int64, c_int64, c_int64]),
    "ax2d_gemm_bf16_wgrad_splits": (c_int, [c_int64, c_int64, c_int64]),
    "ax2d_gemm_bf16_wgrad_workspace": (c_int64, [c_int64, c_int64, c_int64]),
    "ax2d_gemm_bf16_wgrad": (c_int, [C.POINTER(CMat), C.POINTER(CMat), C.POINTER(CMat), c_int64, c_int64, c_int64, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "ax2d_convert": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p]),
    "ax2d_act_bwd_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "ax2d_act_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "ax2d_tick": (c_int, [c_void_p, c_void_p]),
    "ax2d_colsum_workspace": (c_int64, [c_int64, c_int64]),
    "ax2d_colsum": (c_int, [C.POINTER(CMat), c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "ax2d_sqnorm_workspace": (c_int64, [c_int64]),
    "ax2d_sqnorm": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ax2d_clip_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float,
                               c_float, c_float, c_float, c_void_p, c_void_p]),
    "ax2d_clip_adam_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ax2d_pack_weights": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "ax2d_unpack_grads": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "ax2d_weighted_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None

# Optional per-launch timer (ops.KernelTimer) -- when set, every kernel-enqueueing entry point is bracketed by a pair of
# CUDA events on the launching stream (inside a CUDA-graph capture: external event-record nodes, so the durations are
# those of the launches INSIDE the replayed step).  None: the calls go straight to the library.
TIMER_HOOK = [None]
_UNTIMED_MARKS = ("_host_", "_workspace", "_supported", "_splits", "_config", "abi_version", "error_string", "last_error",
                  "launch_count")


class _TimedLib:
    """Thin proxy over the CDLL: same attributes; kernel launches go through ``TIMER_HOOK[0].bracket`` when a timer is set."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("ax2d_") or any(m in name for m in _UNTIMED_MARKS):
            setattr(self, name, fn)
            return fn

        def call(*args):
            t = TIMER_HOOK[0]
            if t is None:
                return fn(*args)
            return t.bracket(name, fn, args)

        setattr(self, name, call)
        return call


def load():
    """Load libax2d.so once; raise loudly when it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the hot path. "
            "Build it with `python aimnet_x2d_b200/build.py` (or __graft_entry__.build()).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ax2d_abi_version() != 1:
        raise RuntimeError("libax2d.so ABI version mismatch; rebuild")
    _lib = _TimedLib(lib)
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load()
        raise RuntimeError(f"{what} failed ({lib.ax2d_error_string(rc).decode()}): {lib.ax2d_last_error().decode()}")
