"""ctypes binding of `libfloodsr_b200.so` (the C ABI declared in `include/floodsr_b200.h`).

The library is built in-tree by `floodsr_b200/build.py` (nvcc, sm_100a).  There is no CPU fallback: a
missing library or a missing CUDA device is an error as soon as an engine is created.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

# FLOODSR_B200_LIB selects another build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = Path(os.environ["FLOODSR_B200_LIB"]).resolve() if os.environ.get("FLOODSR_B200_LIB") else Path(__file__).resolve().parent / "libfloodsr_b200.so"

FSR_OK, FSR_E_INVALID, FSR_E_CUDA, FSR_E_ASSERT, FSR_E_UNSUPPORTED = 0, -1, -2, -3, -4
PREC_FP32, PREC_FP16, PREC_FP32_SIMT = 0, 2, 3
WINDOW_HARD, WINDOW_FEATHER = 0, 1
FLAG_DEPTH_NONFINITE, FLAG_DEM_NONFINITE, FLAG_DEM_FLAT_NONZERO, FLAG_DEPTH_NOT_UNIT, FLAG_DEM_NOT_UNIT = 1, 2, 4, 8, 16
FLAG_PRED_NONFINITE = 32
PROF_CATEGORIES = ("normalize", "lr_conv", "lr_misc", "convt", "head", "invert", "blend")


class TileParams(C.Structure):
    """`fsr_tile_params`."""

    _fields_ = [
        ("max_depth", C.c_float),
        ("depth_denom", C.c_float),
        ("normalize_inputs", C.c_int32),
        ("has_depth_nodata", C.c_int32),
        ("depth_nodata", C.c_float),
        ("depth_nodata_tol", C.c_float),
        ("has_dem_nodata", C.c_int32),
        ("dem_nodata", C.c_float),
        ("dem_nodata_tol", C.c_float),
        ("rank_lo", C.c_int32),
        ("rank_hi", C.c_int32),
        ("gamma", C.c_float),
        ("dem_pct_clip", C.c_float),
        ("has_ref_stats", C.c_int32),
        ("ref_p_clip", C.c_float),
        ("ref_dem_min", C.c_float),
        ("ref_dem_max", C.c_float),
    ]


_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)
_U = C.POINTER(C.c_uint32)
_H = C.c_void_p  # fsr_engine*
_P = C.POINTER(TileParams)

# name -> (restype, argtypes); must list every symbol declared in include/floodsr_b200.h
SIGNATURES = {
    "fsr_abi_version": (C.c_int, []),
    "fsr_last_error": (C.c_char_p, []),
    "fsr_device_count": (C.c_int, []),
    "fsr_create": (C.c_int, [C.c_void_p, C.c_size_t, _F, C.c_size_t, C.c_int, C.c_int, C.POINTER(_H)]),
    "fsr_destroy": (C.c_int, [_H]),
    "fsr_contract": (C.c_int, [_H, _I, _I, _I]),
    "fsr_launch_count": (C.c_int64, [_H]),
    "fsr_macs_per_tile": (C.c_int64, [_H]),
    "fsr_run_tiles": (C.c_int, [_H, _F, _F, C.c_int32, _P, _F, _F, _F, _U]),
    "fsr_run_raster": (
        C.c_int,
        [_H, _F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I, C.c_int32, _I, C.c_int32, _F, _P, _F, _F, _U],
    ),
    "fsr_set_windows": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I, C.c_int32, _I, C.c_int32, _F, C.c_void_p]),
    "fsr_band_geometry": (C.c_int, [_H, C.c_int32, C.c_int32, _I, _I, _I, _I, _I]),
    "fsr_band_run_dev": (
        C.c_int,
        [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "fsr_band_finalize_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fsr_band_finalize_rows_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "fsr_band_host_begin": (C.c_int, [_H, _F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _F, C.c_void_p]),
    "fsr_band_host_end": (C.c_int, [_H, C.c_void_p, C.c_int32, _U]),
    "fsr_fetch_flags": (C.c_int, [_H, C.c_void_p, _U]),
    "fsr_resample_bilinear": (C.c_int, [_H, _F, C.c_int32, C.c_int32, _F, C.c_int32, C.c_int32, C.c_void_p]),
    "fsr_resample_bilinear_dev": (
        C.c_int,
        [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "fsr_stage_normalize": (C.c_int, [_H, _F, _F, C.c_int32, _P, _F, _F, _F, _U]),
    "fsr_stage_forward": (C.c_int, [_H, _F, _F, C.c_int32, _F]),
    "fsr_stage_invert": (C.c_int, [_H, _F, C.c_size_t, C.c_float, C.c_float, _F]),
    "fsr_stage_blend": (
        C.c_int,
        [_H, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I, C.c_int32, _I, C.c_int32, _F, C.c_float, _F],
    ),
    "fsr_profile_enable": (C.c_int, [_H, C.c_int32]),
    "fsr_profile_fetch": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "fsr_debug_tensor_shape": (C.c_int, [_H, C.c_int32, _I, _I, _I]),
    "fsr_debug_read_tensor": (C.c_int, [_H, C.c_int32, C.c_int32, _F]),
    "fsr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "fsr_host_free": (None, [C.c_void_p]),
}

class ResampleParams(C.Structure):
    """fsr_resample_params of include/floodsr_b200.h."""

    _fields_ = [
        ("x_a_dst", C.c_double), ("x_c_dst", C.c_double), ("x_a_src", C.c_double), ("x_c_src", C.c_double),
        ("y_a_dst", C.c_double), ("y_c_dst", C.c_double), ("y_a_src", C.c_double), ("y_c_src", C.c_double),
        ("has_src_nodata", C.c_int32), ("src_nodata", C.c_float), ("dst_fill", C.c_float),
    ]


_lib = None


class EngineLibraryError(RuntimeError):
    """The native library is missing or failed (never silently replaced by a CPU path)."""


def load_library() -> C.CDLL:
    """Load the in-tree shared library and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EngineLibraryError(
            f"{LIB_PATH} not found: build it with `python -m floodsr_b200.build` (nvcc, sm_100a). "
            "floodsr_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return (load_library().fsr_last_error() or b"").decode("utf-8", errors="replace")


def check(code: int) -> None:
    """Turn a negative status into RuntimeError (FSR_E_ASSERT is handled by the callers that know the flags)."""
    if code == FSR_OK:
        return
    msg = last_error()
    if code == FSR_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == FSR_E_INVALID:
        raise ValueError(msg)
    raise EngineLibraryError(f"libfloodsr_b200 error {code}: {msg}")


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_F)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_I)


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked host memory (freed when the array is garbage collected)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.fsr_host_alloc(max(n, 1))
    if not ptr:
        raise EngineLibraryError(f"cudaMallocHost({n}) failed")
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            try:
                lib.fsr_host_free(self.p)
            except Exception:
                pass

    # keep the owner alive as long as any view of the buffer is
    buf._fsr_owner = _Owner(ptr)
    return arr
