"""Host-side window geometry for the B200 engine, API-compatible with the reference's `floodsr/tiling.py`.

Same function names, arguments, assertions and results as `floodsr/tiling.py:7-45`; the results are plain
ints / a float32 vector that the engine uploads once per raster, so the device kernels index windows with
exactly the reference's geometry (bit-exact requirement of SURVEY.md section 8 a12-a14).
"""

from __future__ import annotations

from typing import Iterator

import numpy as np


def build_tile_starts(total_size: int, tile_size: int, stride: int) -> list[int]:
    """Window starts every `stride` pixels plus a forced trailing window flush with the far edge."""
    assert total_size > 0, f"total_size must be > 0; got {total_size}"
    assert tile_size > 0, f"tile_size must be > 0; got {tile_size}"
    assert stride > 0, f"stride must be > 0; got {stride}"
    n_regular = (max(total_size - tile_size + 1, 1) + stride - 1) // stride
    starts = [i * stride for i in range(n_regular)]
    trailing = total_size - tile_size
    if starts[-1] != trailing:
        starts.append(trailing)
    return starts


def iter_window_origins(
    y_starts: list[int], x_starts: list[int], *, use_progress: bool = False, desc: str = "windowed inference"
) -> Iterator[tuple[int, int, int, int]]:
    """Row-major `(yi, xi, y0, x0)`; this order is also the blend kernel's accumulation order."""
    windows = ((yi, xi, y0, x0) for yi, y0 in enumerate(y_starts) for xi, x0 in enumerate(x_starts))
    if use_progress:
        from tqdm import tqdm

        return tqdm(windows, desc=desc, total=len(y_starts) * len(x_starts), unit="window")
    return windows


def build_feather_ramp(tile_size: int, overlap: int) -> np.ndarray:
    """1-D feather weights: linear ramps over `overlap` pixels at both ends, floor 1e-3 (float32)."""
    assert tile_size > 0, f"tile_size must be > 0; got {tile_size}"
    assert overlap >= 0, f"overlap must be >= 0; got {overlap}"
    assert overlap < tile_size, f"overlap must be < tile_size; got overlap={overlap}, tile_size={tile_size}"
    weights = np.ones(tile_size, dtype=np.float32)
    if overlap > 0:
        # identical numpy call to the reference so the float32 values are bit-identical
        inner = np.linspace(0.0, 1.0, overlap + 2, dtype=np.float32)[1:-1]
        weights[:overlap] = inner
        weights[tile_size - overlap :] = inner[::-1]
    return np.clip(weights, 1e-3, 1.0)


def padded_extent(size: int, tile_size: int) -> int:
    """Extent after the reference's zero padding to whole tiles (`ResUNet_16x_DEM.py:215-217`)."""
    return -(-int(size) // int(tile_size)) * int(tile_size)


def window_grid(h: int, w: int, tile_size: int, window_method: str, overlap_hr: int) -> tuple[list[int], list[int]]:
    """Window origins of the tile loop for a padded raster (`ResUNet_16x_DEM.py:299-300, :325-326`)."""
    assert window_method in {"hard", "feather"}, f"unsupported window_method={window_method}"
    ph, pw = padded_extent(h, tile_size), padded_extent(w, tile_size)
    if window_method == "hard":
        return list(range(0, ph, tile_size)), list(range(0, pw, tile_size))
    stride = tile_size - overlap_hr
    if overlap_hr <= 0:
        raise AssertionError("feather windowing requires overlap_lr > 0")
    if stride <= 0:
        raise AssertionError(f"feather stride must be > 0; overlap_hr={overlap_hr}, tile={tile_size}")
    return build_tile_starts(ph, tile_size, stride), build_tile_starts(pw, tile_size, stride)


def split_tile_rows(n_rows: int, n_parts: int) -> list[tuple[int, int]]:
    """Contiguous row-band partition of the window grid's tile rows, larger bands first."""
    assert n_rows > 0 and n_parts > 0
    base, extra = divmod(n_rows, n_parts)
    out, r = [], 0
    for p in range(n_parts):
        n = base + (1 if p < extra else 0)
        out.append((r, r + n))
        r += n
    return out
