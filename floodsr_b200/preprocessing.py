"""Host-side parameter resolution for the device normalisation kernels.

The arithmetic of the reference's `floodsr/preprocessing.py:97-172` runs on the GPU; what stays on the
host is scalar work whose exact float32 results the kernels need as inputs, plus the reference's argument
validation (same AssertionError conditions and wording).
"""

from __future__ import annotations

import numpy as np


def depth_log1p_denom(max_depth: float) -> float:
    """`_depth_log1p_denom` (`preprocessing.py:129-138`)."""
    max_depth = float(max_depth)
    if not np.isfinite(max_depth) or max_depth <= 0:
        raise AssertionError(f"max_depth must be finite and > 0; got {max_depth}")
    denom = float(np.log1p(max_depth))
    if not np.isfinite(denom) or denom <= 0:
        raise AssertionError(f"log1p(max_depth) must be finite and > 0; got {denom}")
    return denom


def check_pct_clip(pct_clip: float) -> float:
    """Validation at `preprocessing.py:108-109`."""
    pct_clip = float(pct_clip)
    if not np.isfinite(pct_clip) or not (0 < pct_clip <= 100):
        raise AssertionError(f"dem_pct_clip must be finite and in (0, 100]; got {pct_clip}")
    return pct_clip


def percentile_ranks(n: int, pct: float) -> tuple[int, int, float]:
    """Order statistics and weight of `np.nanpercentile(x_float32, pct)` (numpy 'linear' method).

    numpy divides the python-float percentile by `float32(100)`, multiplies by `n - 1` and takes the
    fractional part all in float32 (`numpy/lib/_nanfunctions_impl.py` `nanpercentile`,
    `_function_base_impl.py` `_quantile`/`_get_gamma`), so the virtual index is NOT the float64 one.
    Returns `(rank_lo, rank_hi, gamma)`; the kernel evaluates numpy's `_lerp` on the two selected values.
    """
    q = np.float32(pct) / np.float32(100)
    virtual = np.float32(n - 1) * q
    lo = int(np.floor(virtual))
    gamma = float(np.float32(np.float64(virtual) - np.float64(lo)))
    hi = lo + 1
    lo = min(max(lo, 0), n - 1)
    hi = min(max(hi, 0), n - 1)
    return lo, hi, gamma


def nodata_tolerance(nodata: float | None) -> tuple[int, float, float]:
    """`np.isclose(x_float32, nodata)` as a float32 threshold (`preprocessing.py:167-172`).

    isclose tests `abs(x - nodata) <= atol + rtol*abs(nodata)` (atol 1e-8, rtol 1e-5) with the python-float
    right-hand side rounded to float32 by the comparison; non-finite nodata only matches by equality.
    Returns `(has_nodata, nodata_f32, tol_f32)` with tol < 0 meaning equality-only.
    """
    if nodata is None:
        return 0, 0.0, -1.0
    nd = float(nodata)
    if not np.isfinite(nd):
        return 1, nd, -1.0
    nd32 = float(np.float32(nd))
    return 1, nd32, float(np.float32(1e-8 + 1e-5 * abs(nd)))
