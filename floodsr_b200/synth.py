"""Deterministic synthetic inputs: terrain-like DEMs and patchy flood depths.

Used by tests, `bench.py` and `__graft_entry__.smoke()`; there is no network for the reference's
Git-LFS rasters, so every workload is generated from a seed.
"""

from __future__ import annotations

import numpy as np


def synth_dem(h: int, w: int, seed: int = 0, y0: int = 0, x0: int = 0) -> np.ndarray:
    """Smooth terrain (sum of low-frequency sinusoids, ~200-1200 m) plus small noise, float32 [h, w].

    `y0/x0` offset the coordinate grid so that bands of one large raster can be generated independently.
    """
    rng = np.random.default_rng([seed, y0, x0, h, w])
    yy = (np.arange(h, dtype=np.float64) + y0)[:, None]
    xx = (np.arange(w, dtype=np.float64) + x0)[None, :]
    z = (
        700.0
        + 260.0 * np.sin(yy / 811.0 + 0.3 * seed)
        + 170.0 * np.cos(xx / 1237.0 + 0.7 * seed)
        + 60.0 * np.sin((yy + 2.0 * xx) / 173.0)
        + 9.0 * np.cos((3.0 * yy - xx) / 41.0)
    )
    z = z + rng.normal(0.0, 0.05, size=(h, w))
    return z.astype(np.float32)


def synth_depth(h: int, w: int, seed: int = 0, y0: int = 0, x0: int = 0) -> np.ndarray:
    """Low-res water depth: smooth blobs, ~40 % dry, a few cells above max_depth=5 m, float32 [h, w]."""
    rng = np.random.default_rng([seed + 1, y0, x0, h, w])
    yy = (np.arange(h, dtype=np.float64) + y0)[:, None]
    xx = (np.arange(w, dtype=np.float64) + x0)[None, :]
    base = 0.9 * np.sin(yy / 9.0 + seed) * np.cos(xx / 13.0 - seed) + 0.5 * np.sin((yy + xx) / 5.0)
    d = base + rng.normal(0.1, 0.45, size=(h, w))
    d = np.where(rng.random((h, w)) < 0.01, d + 5.5, d)
    return np.clip(d, 0.0, 6.0).astype(np.float32)


def synth_tile(seed: int = 0, lr: int = 32, scale: int = 16) -> tuple[np.ndarray, np.ndarray]:
    """One model tile: (depth_lr [lr, lr], dem_hr [lr*scale, lr*scale])."""
    return synth_depth(lr, lr, seed), synth_dem(lr * scale, lr * scale, seed)


def synth_raster(h: int, w: int, seed: int = 0, scale: int = 16) -> tuple[np.ndarray, np.ndarray]:
    """Model-space raster pair: (depth_lr [h//scale, w//scale], dem_hr [h, w])."""
    return synth_depth(h // scale, w // scale, seed), synth_dem(h, w, seed)
