"""Oracle restatement of the reference's tile loop + mosaic (TEST INFRASTRUCTURE ONLY).

Follows `ModelWorker._run_tiled_model_on_prepared`
(`/root/reference/floodsr/models/ResUNet_16x_DEM.py:140-393`) line block by line block, taking the
prepared model-space arrays directly instead of GeoTIFF paths (the reference re-reads them with
rasterio at :184-185, which is out of scope and unavailable offline).
"""

from __future__ import annotations

import math

import numpy as np

from oracle.tiling_np import feather_ramp, tile_starts, window_origins


def run_tiled(
    engine,
    depth_lr_raw: np.ndarray,
    dem_hr_raw: np.ndarray,
    *,
    max_depth: float = 5.0,
    dem_pct_clip: float = 95.0,
    model_lr_tile: int = 32,
    model_scale: int = 16,
    contract_hr_tile: int = 512,
    window_method: str = "feather",
    overlap_lr: int = 8,
    return_tiles: bool = False,
):
    """Return (prediction_depth_m, n_tiles, tile_dem_stats_summary[, tiles]) exactly as the reference does."""
    assert window_method in {"hard", "feather"}, f"unsupported window_method={window_method}"
    depth_lr_raw = np.asarray(depth_lr_raw, dtype=np.float32)
    dem_hr_raw = np.asarray(dem_hr_raw, dtype=np.float32)
    # :186-189
    assert depth_lr_raw.ndim == 2, f"aligned depth must be 2D; got {depth_lr_raw.shape}"
    assert dem_hr_raw.ndim == 2, f"aligned DEM must be 2D; got {dem_hr_raw.shape}"
    assert np.isfinite(depth_lr_raw).all(), "aligned depth contains non-finite values"
    assert np.isfinite(dem_hr_raw).all(), "aligned DEM contains non-finite values"
    # :193-203
    crop_h, crop_w = dem_hr_raw.shape
    exp_h, exp_w = crop_h // model_scale, crop_w // model_scale
    assert exp_h > 0 and exp_w > 0, f"expected low-resolution shape invalid {(exp_h, exp_w)}"
    assert depth_lr_raw.shape == (exp_h, exp_w), (
        f"depth shape {depth_lr_raw.shape} does not match crop/scale target {(exp_h, exp_w)}"
    )
    # :215-235 zero padding up to a whole number of tiles
    pad_h = int(math.ceil(crop_h / contract_hr_tile)) * contract_hr_tile - crop_h
    pad_w = int(math.ceil(crop_w / contract_hr_tile)) * contract_hr_tile - crop_w
    dem_pad = np.pad(dem_hr_raw, ((0, pad_h), (0, pad_w)), mode="constant", constant_values=0.0)
    hr_pad_h, hr_pad_w = dem_pad.shape
    extra_h = hr_pad_h // model_scale - depth_lr_raw.shape[0]
    extra_w = hr_pad_w // model_scale - depth_lr_raw.shape[1]
    assert extra_h >= 0 and extra_w >= 0, f"computed LR padding must be >= 0; got {(extra_h, extra_w)}"
    depth_pad = np.pad(depth_lr_raw, ((0, extra_h), (0, extra_w)), mode="constant", constant_values=0.0)

    overlap_hr = overlap_lr * model_scale
    cache: dict[tuple[int, int], np.ndarray] = {}
    stats_l: list[tuple[float, float, float]] = []

    def predict(y0: int, x0: int) -> np.ndarray:
        """:250-294 tile extraction, engine call, cache by origin."""
        key = (int(y0), int(x0))
        if key in cache:
            return cache[key]
        ly, lx = y0 // model_scale, x0 // model_scale
        depth_tile = depth_pad[ly : ly + model_lr_tile, lx : lx + model_lr_tile]
        dem_tile = dem_pad[y0 : y0 + contract_hr_tile, x0 : x0 + contract_hr_tile]
        assert depth_tile.shape == (model_lr_tile, model_lr_tile)
        assert dem_tile.shape == (contract_hr_tile, contract_hr_tile)
        res = engine.run_tile(
            depth_tile,
            dem_tile,
            max_depth=max_depth,
            dem_pct_clip=dem_pct_clip,
            dem_ref_stats=None,
            normalize_inputs=True,
            depth_lr_nodata=None,
            dem_hr_nodata=None,
        )
        pred = res["prediction_m"]
        assert pred.shape == (contract_hr_tile, contract_hr_tile)
        st = res.get("dem_stats_used")
        if isinstance(st, dict):
            stats_l.append((float(st.get("p_clip", 0.0)), float(st.get("dem_min", 0.0)), float(st.get("dem_max", 0.0))))
        cache[key] = pred
        return pred

    if window_method == "hard":
        # :297-314
        sr_pad = np.zeros_like(dem_pad, dtype=np.float32)
        ys = list(range(0, hr_pad_h, contract_hr_tile))
        xs = list(range(0, hr_pad_w, contract_hr_tile))
        for _yi, _xi, y0, x0 in window_origins(ys, xs):
            sr_pad[y0 : y0 + contract_hr_tile, x0 : x0 + contract_hr_tile] = predict(y0, x0)
    else:
        # :315-363
        stride_hr = contract_hr_tile - overlap_hr
        if overlap_lr <= 0:
            raise AssertionError("feather windowing requires overlap_lr > 0")
        if stride_hr <= 0:
            raise AssertionError(f"feather stride must be > 0; overlap_lr={overlap_lr}, tile={contract_hr_tile}")
        ys = tile_starts(hr_pad_h, contract_hr_tile, stride_hr)
        xs = tile_starts(hr_pad_w, contract_hr_tile, stride_hr)
        ramp = feather_ramp(contract_hr_tile, overlap_hr)
        accum = np.zeros_like(dem_pad, dtype=np.float32)
        wsum = np.zeros_like(dem_pad, dtype=np.float32)
        for yi, xi, y0, x0 in window_origins(ys, xs):
            pred = predict(y0, x0)
            wy = ramp.copy()
            wx = ramp.copy()
            if overlap_hr > 0:
                if yi == 0:
                    wy[:overlap_hr] = 1.0
                if yi == len(ys) - 1:
                    wy[-overlap_hr:] = 1.0
                if xi == 0:
                    wx[:overlap_hr] = 1.0
                if xi == len(xs) - 1:
                    wx[-overlap_hr:] = 1.0
            weight = np.outer(wy, wx).astype(np.float32, copy=False)
            accum[y0 : y0 + contract_hr_tile, x0 : x0 + contract_hr_tile] += pred * weight
            wsum[y0 : y0 + contract_hr_tile, x0 : x0 + contract_hr_tile] += weight
        sr_pad = np.divide(accum, np.maximum(wsum, 1e-6), out=np.zeros_like(accum), where=wsum > 0)

    # :366-389
    summary = None
    if stats_l:
        st = np.asarray(stats_l, dtype=np.float32)
        rng = st[:, 2] - st[:, 1]
        summary = {
            "tile_count": float(st.shape[0]),
            "dem_p_clip_min": float(st[:, 0].min()),
            "dem_p_clip_mean": float(st[:, 0].mean()),
            "dem_p_clip_max": float(st[:, 0].max()),
            "dem_range_min": float(rng.min()),
            "dem_range_mean": float(rng.mean()),
            "dem_range_max": float(rng.max()),
        }
    # :391
    out = np.clip(sr_pad[:crop_h, :crop_w], 0.0, max_depth).astype(np.float32, copy=False)
    if return_tiles:
        return out, len(cache), summary, {"origins": (ys, xs), "cache": cache}
    return out, len(cache), summary
