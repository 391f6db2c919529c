"""ctypes binding of librnerf_b200.so (the C ABI declared in include/rnerf_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch only owns the device buffers and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# RNERF_LIB: development aid for A/B runs of two builds of the same sources (never a different implementation)
LIB_PATH = os.environ.get("RNERF_LIB") or os.path.join(_HERE, "librnerf_b200.so")

_lib: Optional[C.CDLL] = None
ABI_VERSION = 11      # include/rnerf_b200.h RNERF_ABI_VERSION

c_f32p = C.c_void_p
c_i64 = C.c_int64
Int3 = C.c_int * 3
Dbl3 = C.c_double * 3

# name -> (restype, argtypes); must list every symbol of include/rnerf_b200.h
SIGNATURES = {
    "rnerf_abi_version": (C.c_int, []),
    "rnerf_last_error": (C.c_char_p, []),
    "rnerf_launch_count": (C.c_uint64, []),
    "rnerf_grid_blur": (C.c_int, [c_f32p, c_f32p, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_void_p]),
    "rnerf_grid_table": (C.c_int, [c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                   C.c_void_p]),
    "rnerf_grid_lookup": (C.c_int, [c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                    c_i64, c_f32p, C.c_void_p]),
    "rnerf_grid_brick_count": (C.c_int64, [C.POINTER(C.c_int)]),
    "rnerf_grid_bricks": (C.c_int, [c_f32p, C.POINTER(C.c_int), c_f32p, C.c_void_p]),
    "rnerf_march_fwd": (C.c_int, [c_f32p, c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                  c_f32p, c_i64, C.c_double, C.c_double, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_so3_weight_floats": (C.c_size_t, []),
    "rnerf_so3_saved_floats": (C.c_size_t, [c_i64, C.c_int]),
    "rnerf_march_all_fwd": (C.c_int, [c_f32p, c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                      c_f32p, c_i64, C.c_double, C.c_double, C.c_int, C.c_int, c_f32p, C.POINTER(C.c_double),
                                      c_f32p, C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_path_dirs": (C.c_int, [c_f32p, C.c_int, c_i64, C.c_int, c_f32p, C.c_void_p]),
    "rnerf_select": (C.c_int, [c_f32p, C.c_int, c_i64, C.c_int, C.c_void_p, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_encmlp_packed_bytes": (C.c_size_t, []),
    "rnerf_encmlp_pack": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "rnerf_encmlp_fwd": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p]),
    "rnerf_encmlp_fwd_debug": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p, C.c_void_p]),
    "rnerf_encmlp_fwd_profile": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p, C.c_void_p]),
    "rnerf_encmlp_fwd_train": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rnerf_mlp_dgrad_packed_bytes": (C.c_size_t, []),
    "rnerf_mlp_dgrad_pack": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "rnerf_mlp_dgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_f32p, c_i64, C.c_void_p, C.c_void_p]),
    "rnerf_mlp_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, c_i64, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_mlp_wgrad_batched": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_int), c_i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.c_void_p]),
    "rnerf_mlp_head_grad": (C.c_int, [C.c_void_p, c_f32p, c_i64, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_generate_rays": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_radiance_loss_ws_floats": (C.c_size_t, []),
    "rnerf_radiance_loss_fwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_f32p, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_double, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_radiance_loss_bwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_f32p, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_double, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_sq_err": (C.c_int, [c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p]),
    "rnerf_sumsq": (C.c_int, [c_f32p, c_i64, c_f32p, C.c_void_p]),
    "rnerf_grad_sumsq": (C.c_int, [c_f32p, c_f32p, c_i64, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_adam_step": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_bkgd_weight_floats": (C.c_size_t, []),
    "rnerf_bkgd_mlp_fwd": (C.c_int, [c_f32p, c_f32p, c_i64, c_i64, c_f32p, C.c_void_p]),
    "rnerf_bkgd_mlp_bwd": (C.c_int, [c_f32p, c_f32p, c_i64, c_i64, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_bkgd_mlp_bwd_dirs": (C.c_int, [c_f32p, c_f32p, c_i64, c_i64, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_mlp_input_grad_packed_floats": (C.c_size_t, []),
    "rnerf_mlp_input_grad_pack": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_mlp_input_grad": (C.c_int, [C.c_void_p, c_i64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_so3_predict": (C.c_int, [c_f32p, C.POINTER(C.c_double), c_f32p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p]),
    "rnerf_so3_tc_packed_bytes": (C.c_size_t, []),
    "rnerf_so3_tc_pack": (C.c_int, [c_f32p, C.c_void_p, C.c_void_p]),
    "rnerf_so3_predict_tc": (C.c_int, [C.c_void_p, c_f32p, C.POINTER(C.c_double), c_f32p, c_f32p, c_f32p, c_i64, c_f32p, C.c_void_p]),
    "rnerf_bkgd_mlp_fwd_tc": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64, c_i64, c_f32p, C.c_void_p]),
    "rnerf_so3_transposed_floats": (C.c_size_t, []),
    "rnerf_so3_transpose": (C.c_int, [c_f32p, c_f32p, C.c_void_p]),
    "rnerf_march_all_bwd": (C.c_int, [c_f32p, c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                      C.c_int, c_i64, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, c_f32p, c_f32p,
                                      c_f32p, c_f32p, C.POINTER(C.c_double), c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_grid_table_bwd": (C.c_int, [c_f32p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                       C.c_void_p]),
    "rnerf_composite_fwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i64, C.c_int, C.c_int, C.c_double,
                                      C.c_double, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_composite_bwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i64, C.c_int, C.c_int, C.c_double,
                                      C.c_double, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_resample": (C.c_int, [c_f32p, C.c_int, c_f32p, c_i64, C.c_int, c_f32p, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p,
                                 c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "rnerf_bbox_tail_mask": (C.c_int, [c_f32p, c_i64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), c_f32p,
                                       c_f32p, C.c_void_p]),
}


class RnerfError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RnerfError(
            f"{LIB_PATH} is missing: build it with `python -m samplenerfro_b200.build` "
            "(there is no CPU fallback for the rendering path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rnerf_abi_version() != ABI_VERSION:
        raise RnerfError(f"ABI version mismatch: library reports {lib.rnerf_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rnerf_last_error().decode("utf-8", "replace")
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise RnerfError(f"{what}: {kind} {rc}: {msg}")


def launch_count() -> int:
    return int(load().rnerf_launch_count())
