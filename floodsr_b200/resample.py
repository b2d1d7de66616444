"""Grid changes around the tile loop on the GPU (SURVEY.md section 8 f#2).

Mirrors the two `rasterio.warp.reproject(..., resampling=Resampling.bilinear)` calls of the reference:
  * `align_dem_to_model_grid`  floodsr/preprocessing.py:367-398: DEM clipped to the depth raster's bounds -> model grid of
    shape (LR rows * scale, LR cols * scale) spanning the same bounds, nodata-aware, then nodata -> 0;
  * `prediction_to_raw_grid`   floodsr/models/ResUNet_16x_DEM.py:552-573: model-grid prediction -> the raw DEM grid.
Transforms are 6-tuples in rasterio.Affine order (a, b, c, d, e, f); only north-up grids (b = d = 0) in one CRS are
handled, which is what the reference asserts before it gets here (preprocessing.py:309-323).  The arithmetic runs in
libfloodsr_b200 (`fsr_resample_bilinear`); there is no CPU fallback.  GDAL is not available offline, so parity with GDAL's
own kernel is unpinned: the restated algorithm and its source are described in oracle/resample_np.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from floodsr_b200 import _lib


def bounds_to_transform(west: float, south: float, east: float, north: float, width: int, height: int):
    """`rasterio.transform.from_bounds` (preprocessing.py:293, :374): translation(west, north) * scale(dx, -dy)."""
    return ((east - west) / width, 0.0, west, 0.0, (south - north) / height, north)


def _params(src_transform, dst_transform, src_nodata, dst_nodata) -> _lib.ResampleParams:
    sa, sb, sc, sd, se, sf = (float(v) for v in tuple(src_transform)[:6])
    da, db, dc, dd, de, df = (float(v) for v in tuple(dst_transform)[:6])
    assert sb == 0.0 and sd == 0.0 and db == 0.0 and dd == 0.0, "rotated grids are not supported"
    p = _lib.ResampleParams()
    p.x_a_dst, p.x_c_dst, p.x_a_src, p.x_c_src = da, dc, sa, sc
    p.y_a_dst, p.y_c_dst, p.y_a_src, p.y_c_src = de, df, se, sf
    p.has_src_nodata = 0 if src_nodata is None else 1
    p.src_nodata = 0.0 if src_nodata is None else float(np.float32(src_nodata))
    p.dst_fill = 0.0 if dst_nodata is None else float(np.float32(dst_nodata))
    return p


def resample_bilinear(engine, src: np.ndarray, src_transform, dst_shape, dst_transform, src_nodata=None, dst_nodata=None) -> np.ndarray:
    """`reproject(source=src, destination=empty(dst_shape), ..., resampling=bilinear)` for north-up grids, on `engine`'s GPU."""
    src = np.ascontiguousarray(src, dtype=np.float32)
    assert src.ndim == 2, f"source must be 2-D; got {src.shape}"
    dh, dw = (int(v) for v in dst_shape)
    assert dh > 0 and dw > 0, f"destination shape invalid {(dh, dw)}"
    if getattr(engine, "_handle", None) is None:
        engine.load()
    dst = np.empty((dh, dw), dtype=np.float32)
    p = _params(src_transform, dst_transform, src_nodata, dst_nodata)
    lib = _lib.load_library()
    _lib.check(lib.fsr_resample_bilinear(engine._handle, _lib.fptr(src), src.shape[0], src.shape[1], _lib.fptr(dst), dh, dw, C.byref(p)))
    return dst


def align_dem_to_model_grid(engine, dem_crop: np.ndarray, dem_crop_transform, depth_bounds, depth_shape, scale: int, dem_nodata=None) -> dict:
    """The resampling half of `align_inputs_for_model_scale` (preprocessing.py:367-398), arrays in, arrays out."""
    target_h, target_w = int(depth_shape[0] * scale), int(depth_shape[1] * scale)
    assert target_h > 0 and target_w > 0, f"target HR shape invalid {(target_h, target_w)}"
    dem_model_transform = bounds_to_transform(*depth_bounds, width=target_w, height=target_h)
    dem_model = resample_bilinear(engine, dem_crop, dem_crop_transform, (target_h, target_w), dem_model_transform, dem_nodata, dem_nodata)
    if dem_nodata is not None:  # replace_nodata_with_zero (preprocessing.py:167-172, :388)
        dem_model = np.where(np.isclose(dem_model, dem_nodata), 0.0, dem_model).astype(np.float32, copy=False)
    if not np.isfinite(dem_model).all():
        raise AssertionError("resampled DEM contains non-finite values")
    ct = tuple(dem_crop_transform)[:6]
    was_resampled = bool(
        dem_model.shape != tuple(dem_crop.shape) or not all(np.isclose((dem_model_transform[0], dem_model_transform[4]), (ct[0], ct[4])))
    )
    return {"dem_hr": dem_model, "dem_hr_transform": dem_model_transform, "crop_shape": (target_h, target_w), "resampled": was_resampled}


def prediction_to_raw_grid(engine, prediction_model_m: np.ndarray, model_transform, raw_shape, raw_transform) -> np.ndarray:
    """Post-resampling of the worker (ResUNet_16x_DEM.py:552-573): skipped when the shapes already agree."""
    pred = np.asarray(prediction_model_m, dtype=np.float32)
    if tuple(raw_shape) == tuple(pred.shape):
        return pred
    return resample_bilinear(engine, pred, model_transform, raw_shape, raw_transform)
