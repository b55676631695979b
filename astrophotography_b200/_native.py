"""ctypes binding of ``libapgpu.so`` (the C ABI declared in ``include/apgpu.h``).

The library is built in-tree by ``astrophotography_b200.build`` (nvcc, sm_100a).
There is no CPU fallback: every wrapper raises ``RuntimeError`` when the
library is missing, when no CUDA device is present, or when the native call
reports an error (message from ``apgpu_last_error()``), mirroring the
reference's RuntimeError convention (core/ApCalibrate.py:123-125).
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libapgpu.so")

# (name, restype, argtypes) -- one row per symbol declared in include/apgpu.h
_c = ctypes
_P = _c.c_void_p
SIGNATURES = {
    "apgpu_abi_version": (_c.c_int, []),
    "apgpu_last_error": (_c.c_char_p, []),
    "apgpu_launch_count": (_c.c_uint64, []),
    "apgpu_stack_reduce_f32": (_c.c_int, [
        _c.POINTER(_P), _c.c_int, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int,
        _c.c_double, _c.c_double, _c.c_int, _c.c_int, _c.c_int,
        _P, _c.c_int, _P, _c.c_int, _P, _P, _c.c_int, _P]),
    "apgpu_stack_reduce_u16": (_c.c_int, [
        _c.POINTER(_P), _c.c_int, _c.c_int, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int,
        _c.c_double, _c.c_double, _c.c_int, _c.c_int, _c.c_int,
        _P, _c.c_int, _P, _c.c_int, _P, _P, _c.c_int, _P]),
    "apgpu_stack_kernel_name": (_c.c_char_p, [
        _c.c_int, _c.c_int, _c.c_double, _c.c_double, _c.c_int, _c.c_int, _c.c_int,
        _c.c_int, _c.c_int, _c.c_int]),
    "apgpu_stack_last_staging": (_c.c_int, []),
    "apgpu_flat_norm_workspace_bytes": (_c.c_size_t, [_c.c_int64]),
    "apgpu_flat_norm_f32": (_c.c_int, [_P, _c.c_int64, _P, _c.c_size_t, _P, _P]),
    "apgpu_flat_divide_f32": (_c.c_int, [_P, _P, _P, _c.c_int64, _P]),
    "apgpu_calibrate_f32": (_c.c_int, [_P, _P, _P, _P, _c.c_float, _c.c_int, _P, _c.c_int64, _P]),
    "apgpu_calibrate_u16": (_c.c_int, [_P, _c.c_float, _c.c_int, _P, _P, _P, _c.c_float, _c.c_int,
                                       _P, _c.c_int64, _P]),
    "apgpu_fix_badpix_f32": (_c.c_int, [_P, _P, _c.c_int, _c.c_int64, _c.c_int64, _c.c_int64,
                                        _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int, _c.c_int,
                                        _P, _P, _P]),
    "apgpu_calibrate_repair": (_c.c_int, [_P, _c.c_int, _c.c_float, _c.c_int, _P, _P, _P, _c.c_float, _c.c_int,
                                          _P, _c.c_int64, _c.c_int64, _c.c_int, _c.c_int, _P, _c.c_int, _P, _P]),
    "apgpu_imarith_f32": (_c.c_int, [_P, _P, _c.c_int, _c.c_double, _c.c_int, _P, _c.c_int64, _P]),
    "apgpu_image_stats_workspace_bytes": (_c.c_size_t, [_c.c_int64]),
    "apgpu_sigma_clipped_stats_f32": (_c.c_int, [_P, _c.c_int64, _c.c_double, _c.c_int, _P,
                                                 _c.c_size_t, _P, _P]),
    "apgpu_threshold_mask_f32": (_c.c_int, [_P, _c.c_int64, _c.c_double, _c.c_double, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def library_path() -> str:
    return LIB_PATH


def load():
    """Load (once) and return the ctypes handle; RuntimeError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"native CUDA library {LIB_PATH} not found: run "
                "`python -m astrophotography_b200.build` (there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                continue           # optional symbols are checked by tests against the header
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "apgpu"):
    if status != 0:
        msg = load().apgpu_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(load().apgpu_launch_count())


def require_cuda():
    """torch device plumbing for the host classes; fails loudly without a GPU."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("astrophotography_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    load()
    return torch
