"""Oracle restatement of the reference's window geometry (TEST INFRASTRUCTURE ONLY).

Follows `/root/reference/floodsr/tiling.py:7-45`.  Geometry must be reproduced bit-exactly.
"""

from __future__ import annotations

import numpy as np


def tile_starts(total_size: int, tile_size: int, stride: int) -> list[int]:
    """tiling.py:7-16: regular starts every `stride`, plus a forced last start at `total - tile`."""
    assert total_size > 0, f"total_size must be > 0; got {total_size}"
    assert tile_size > 0, f"tile_size must be > 0; got {tile_size}"
    assert stride > 0, f"stride must be > 0; got {stride}"
    starts = []
    s = 0
    limit = max(total_size - tile_size + 1, 1)
    while s < limit:
        starts.append(s)
        s += stride
    if starts[-1] != total_size - tile_size:
        starts.append(total_size - tile_size)
    return starts


def window_origins(y_starts, x_starts):
    """tiling.py:19-31 without the tqdm wrapper: row-major (yi, xi, y0, x0)."""
    for yi, y0 in enumerate(y_starts):
        for xi, x0 in enumerate(x_starts):
            yield yi, xi, y0, x0


def feather_ramp(tile_size: int, overlap: int) -> np.ndarray:
    """tiling.py:34-45: ones with linspace(0,1,overlap+2)[1:-1] ramps on both ends, clipped to [1e-3, 1]."""
    assert tile_size > 0, f"tile_size must be > 0; got {tile_size}"
    assert overlap >= 0, f"overlap must be >= 0; got {overlap}"
    assert overlap < tile_size, f"overlap must be < tile_size; got overlap={overlap}, tile_size={tile_size}"
    w = np.ones(tile_size, dtype=np.float32)
    if overlap > 0:
        ramp = np.linspace(0.0, 1.0, overlap + 2, dtype=np.float32)[1:-1]
        w[:overlap] = ramp
        w[-overlap:] = ramp[::-1]
    return np.clip(w, 1e-3, 1.0)
