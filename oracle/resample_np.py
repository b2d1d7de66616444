"""CPU restatement of the bilinear grid change the reference delegates to GDAL (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(floodsr_b200/resample.py -> fsr_resample_bilinear) never does.

** PARITY UNPINNED ** The arithmetic lives in a third-party dependency that is absent from /root/reference and from this
container: rasterio.warp.reproject -> GDAL's warp kernel (rasterio 1.5.0 / libgdal 3.12.2 in the reference's deploy
lock file, container/miniforge/conda-env-deploy.lock.yml:57,118).  The reference's call sites are
  * floodsr/preprocessing.py:371-387   raw DEM grid -> model grid (LR shape x 16), src_nodata = dst_nodata = DEM nodata,
  * floodsr/models/ResUNet_16x_DEM.py:554-573   prediction on the model grid -> raw DEM grid, no nodata,
both Resampling.bilinear, same CRS, north-up affine transforms.  No golden raster for either call exists in the
reference's tests (tests/data rasters are LFS pointers), so this file restates GDAL's published algorithm
(alg/gdalwarpkernel.cpp, from memory of the 3.x sources) and nothing pins it against GDAL itself:

  * the centre of destination pixel (i, j) is mapped through the two geotransforms to fractional source pixel
    coordinates (sx, sy); destinations whose centre falls outside the source raster keep the destination nodata
    (rasterio initialises the destination with dst_nodata, 0 when none is given);
  * scale = destination pixels per source pixel along each axis.  With both scales >= 0.95 GDAL uses the 4-sample
    formula (GWKBilinearResample4Sample): base = floor(s - 0.5), ratio = 1.5 - (s - base), weights ratio / 1 - ratio
    per axis, neighbours outside the raster or equal to src_nodata are skipped and the sum is divided by the weight
    actually used (value dropped when that weight is < 1e-5);
  * otherwise (down-sampling, e.g. 4096 -> 3840 = 0.9375) the general kernel (GWKResample): the triangle filter is
    widened by 1 / scale: radius = ceil(1 / scale), taps i = -radius .. radius around base = floor(s - 0.5) with
    weight max(0, 1 - |(i - (s - 0.5 - base)) * scale|), same skipping and normalisation (threshold 1e-6);
  * accumulation in float64, result cast to float32.
"""
from __future__ import annotations

import math

import numpy as np


def _axis_coords(n_dst: int, a_dst: float, c_dst: float, a_src: float, c_src: float) -> np.ndarray:
    """Fractional source pixel coordinate of every destination pixel centre along one axis."""
    i = np.arange(n_dst, dtype=np.float64) + 0.5
    return ((a_dst * i + c_dst) - c_src) / a_src


def _axis_taps(s: np.ndarray, n_src: int, four: bool, scale: float):
    """Per destination coordinate: tap indices [n_dst, k], weights [n_dst, k] (0 where the tap is outside the raster)."""
    base = np.floor(s - 0.5)
    if four:
        ratio = 1.5 - (s - base)
        idx = np.stack([base, base + 1], axis=1)
        w = np.stack([ratio, 1.0 - ratio], axis=1)
    else:
        k = min(scale, 1.0)  # the filter is only ever widened
        radius = int(math.ceil(1.0 / k))
        delta = (s - 0.5) - base
        offs = np.arange(-radius, radius + 1, dtype=np.float64)
        idx = base[:, None] + offs[None, :]
        w = np.maximum(0.0, 1.0 - np.abs((offs[None, :] - delta[:, None]) * k))
    inside = (idx >= 0) & (idx < n_src)
    w = np.where(inside, w, 0.0)
    idx = np.clip(idx, 0, n_src - 1).astype(np.int64)
    return idx, w


def resample_bilinear(src: np.ndarray, src_transform, dst_shape, dst_transform, src_nodata=None, dst_nodata=None) -> np.ndarray:
    """`rasterio.warp.reproject(..., resampling=bilinear)` for two north-up grids in the same CRS.

    transforms: (a, b, c, d, e, f) with x = a*col + b*row + c, y = d*col + e*row + f (rasterio.Affine order), b = d = 0.
    """
    src = np.asarray(src, dtype=np.float32)
    sa, sb, sc, sd, se, sf = (float(v) for v in tuple(src_transform)[:6])
    da, db, dc, dd, de, df = (float(v) for v in tuple(dst_transform)[:6])
    assert sb == 0.0 and sd == 0.0 and db == 0.0 and dd == 0.0, "rotated grids are not supported"
    dh, dw = (int(v) for v in dst_shape)
    sh, sw = src.shape
    sx = _axis_coords(dw, da, dc, sa, sc)
    sy = _axis_coords(dh, de, df, se, sf)
    x_scale, y_scale = abs(sa / da), abs(se / de)
    four = x_scale >= 0.95 and y_scale >= 0.95  # GDAL switches both axes to the general kernel together
    ix, wx = _axis_taps(sx, sw, four, x_scale)
    iy, wy = _axis_taps(sy, sh, four, y_scale)
    eps = 1e-5 if four else 1e-6
    fill = np.float32(0.0 if dst_nodata is None else dst_nodata)
    out = np.full((dh, dw), fill, dtype=np.float32)
    centre_in_x = (sx >= 0) & (sx < sw)
    centre_in_y = (sy >= 0) & (sy < sh)
    src64 = src.astype(np.float64)
    valid = np.ones_like(src, dtype=bool) if src_nodata is None else (src != np.float32(src_nodata))
    for j in range(dh):  # row by row: bounded memory, fixed accumulation order (y taps outer, x taps inner)
        if not centre_in_y[j]:
            continue
        acc = np.zeros(dw, dtype=np.float64)
        wsum = np.zeros(dw, dtype=np.float64)
        for ky in range(iy.shape[1]):
            if wy[j, ky] == 0.0:
                continue
            row = src64[iy[j, ky]]
            vrow = valid[iy[j, ky]]
            for kx in range(ix.shape[1]):
                ok = vrow[ix[:, kx]] & (wx[:, kx] != 0.0)
                w = np.where(ok, wx[:, kx] * wy[j, ky], 0.0)
                acc = acc + row[ix[:, kx]] * w
                wsum = wsum + w
        good = centre_in_x & (wsum >= eps)
        with np.errstate(invalid="ignore", divide="ignore"):
            val = np.where(wsum == 1.0, acc, acc / wsum)
        out[j] = np.where(good, val.astype(np.float32), fill)
    return out
