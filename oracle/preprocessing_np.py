"""Oracle restatement of the reference's per-tile normalisation (TEST INFRASTRUCTURE ONLY).

Follows `/root/reference/floodsr/preprocessing.py:12-172` function by function; see each docstring for
the exact lines.  All arithmetic is numpy float32 with python-float scalars, as in the reference.
"""

from __future__ import annotations

import numpy as np


def check_numeric(arr, name: str, allow_ranks=None, min_rank: int = 1) -> np.ndarray:
    """`_as_numeric_np_array` (preprocessing.py:12-36): dtype, rank and finiteness validation."""
    out = np.asarray(arr)
    if out.dtype == np.bool_ or not np.issubdtype(out.dtype, np.number):
        raise AssertionError(f"{name} must have numeric dtype; got {out.dtype}")
    if allow_ranks is not None:
        if out.ndim not in allow_ranks:
            raise AssertionError(f"{name} rank must be one of {allow_ranks}; got rank {out.ndim} shape {out.shape}")
        if out.ndim >= 3 and out.shape[-1] != 1:
            raise AssertionError(f"{name} last dim must be 1 for rank >=3; got shape {out.shape}")
    elif out.ndim < min_rank:
        raise AssertionError(f"{name} rank must be >= {min_rank}; got rank {out.ndim} shape {out.shape}")
    if not np.all(np.isfinite(out)):
        raise AssertionError(f"{name} must contain only finite values")
    return out


def replace_nodata_with_zero(arr, nodata):
    """preprocessing.py:167-172: float32 cast, then isclose(nodata) -> 0."""
    a = np.asarray(arr, dtype=np.float32)
    if nodata is None:
        return a
    return np.where(np.isclose(a, nodata), 0.0, a).astype(np.float32, copy=False)


def depth_log1p_denom(max_depth: float) -> float:
    """preprocessing.py:129-138."""
    max_depth = float(max_depth)
    if not np.isfinite(max_depth) or max_depth <= 0:
        raise AssertionError(f"max_depth must be finite and > 0; got {max_depth}")
    denom = float(np.log1p(max_depth))
    if not np.isfinite(denom) or denom <= 0:
        raise AssertionError(f"log1p(max_depth) must be finite and > 0; got {denom}")
    return denom


def scale_depth_log1p(arr, max_depth: float) -> np.ndarray:
    """preprocessing.py:141-151: clip to [0, D], log1p, divide by log1p(D), clip to [0, 1]."""
    denom = depth_log1p_denom(max_depth)
    a = check_numeric(arr, "depth_arr").astype(np.float32, copy=False)
    a = np.clip(a, 0.0, float(max_depth))
    s = np.log1p(a) / denom
    return np.clip(s, 0.0, 1.0).astype(np.float32, copy=False)


def invert_depth_log1p(arr, max_depth: float) -> np.ndarray:
    """preprocessing.py:154-164: clip to [0, 1], expm1(y * log1p(D)), clip to [0, D]."""
    denom = depth_log1p_denom(max_depth)
    a = check_numeric(arr, "normalized_depth_arr").astype(np.float32, copy=False)
    a = np.clip(a, 0.0, 1.0)
    inv = np.expm1(a * denom)
    return np.clip(inv, 0.0, float(max_depth)).astype(np.float32, copy=False)


def parse_dem_stats(ref_stats) -> tuple[float, float, float]:
    """`_parse_dem_normalization_stats` (preprocessing.py:39-58)."""
    missing = {"p_clip", "dem_min", "dem_max"}.difference(ref_stats.keys())
    if missing:
        raise AssertionError(f"DEM ref_stats missing keys: {sorted(missing)}")
    p_clip, dem_min, dem_max = (float(ref_stats[k]) for k in ("p_clip", "dem_min", "dem_max"))
    if not (np.isfinite(p_clip) and np.isfinite(dem_min) and np.isfinite(dem_max)):
        raise AssertionError("DEM ref_stats values must be finite")
    if p_clip < 0:
        raise AssertionError(f"DEM p_clip must be >= 0; got {p_clip}")
    if dem_min > dem_max:
        raise AssertionError(f"DEM dem_min must be <= dem_max; got min={dem_min} max={dem_max}")
    if (dem_max - dem_min) <= 0:
        raise AssertionError(f"DEM range must be > 0; got min={dem_min}, max={dem_max}")
    return p_clip, dem_min, dem_max


def normalize_dem_with_stats(arr, p_clip: float, dem_min: float, dem_max: float) -> np.ndarray:
    """preprocessing.py:61-94: clip to [0, p_clip], min-max scale, clip to [0, 1]; flat all-zero tile -> zeros."""
    if not (np.isfinite(p_clip) and np.isfinite(dem_min) and np.isfinite(dem_max)):
        raise AssertionError("p_clip, dem_min, and dem_max must be finite")
    dem_range = dem_max - dem_min
    if dem_range <= 0:
        if np.isclose(dem_range, 0.0) and np.isclose(dem_min, 0.0):
            a = check_numeric(arr, "dem_arr", allow_ranks=(2, 3, 4)).astype(np.float32, copy=False)
            return np.zeros_like(a)
        raise AssertionError(f"DEM range must be > 0; got min={dem_min}, max={dem_max}")
    a = check_numeric(arr, "dem_arr", allow_ranks=(2, 3, 4)).astype(np.float32, copy=False)
    clipped = np.clip(a, 0.0, float(p_clip))
    norm = (clipped - float(dem_min)) / float(dem_range)
    return np.clip(norm, 0.0, 1.0).astype(np.float32, copy=False)


def dem_tile_stats(arr, pct_clip: float) -> tuple[float, float, float]:
    """Tile-local stats branch of `normalize_dem` (preprocessing.py:106-121)."""
    pct_clip = float(pct_clip)
    if not np.isfinite(pct_clip) or not (0 < pct_clip <= 100):
        raise AssertionError(f"dem_pct_clip must be finite and in (0, 100]; got {pct_clip}")
    a = check_numeric(arr, "dem_arr", allow_ranks=(2, 3, 4)).astype(np.float32, copy=False)
    a = np.clip(a, 0.0, None)
    p_clip = float(np.nanpercentile(a, pct_clip))
    for_stats = np.clip(a, 0.0, p_clip)
    return p_clip, float(np.nanmin(for_stats)), float(np.nanmax(for_stats))


def normalize_dem(arr, pct_clip: float = 95.0, ref_stats=None):
    """preprocessing.py:97-126."""
    if arr is None:
        return None, None
    if ref_stats is None:
        p_clip, dem_min, dem_max = dem_tile_stats(arr, pct_clip)
    else:
        p_clip, dem_min, dem_max = parse_dem_stats(ref_stats)
    norm = normalize_dem_with_stats(arr, p_clip=p_clip, dem_min=dem_min, dem_max=dem_max)
    return norm, {"p_clip": p_clip, "dem_min": dem_min, "dem_max": dem_max}
