"""B200-aware worker: host logic on CPU, array pipeline on the GPU against the oracle."""

from __future__ import annotations

import json
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from floodsr_b200.synth import synth_raster
from floodsr_b200.worker import postprocess_depth, reconcile_tiling, resolve_preprocess_config

CONTRACT = SimpleNamespace(scale=16, depth_lr_hwc=(32, 32, 1), dem_hr_hwc=(512, 512, 1))


def test_resolve_preprocess_config_matches_reference(tmp_path):
    model = tmp_path / "model_infer.onnx"
    model.write_bytes(b"x")
    assert resolve_preprocess_config(model)["max_depth"] == 5.0 and resolve_preprocess_config(model)["dem_pct_clip"] == 95.0
    (tmp_path / "train_config.json").write_text(json.dumps({
        "max_depth": 4.0, "dem_pct_clip": 90.0, "dem_stats": {"p_clip": 3.0, "dem_min": 1.0, "dem_max": 2.0},
        "input_shape": [32, 32, 1], "upscale": 16, "dem_fp": "data/tiles_02dem/x.tif"}))
    mine = resolve_preprocess_config(model, dem_pct_clip=99.0)
    assert mine["max_depth"] == 4.0 and mine["dem_pct_clip"] == 99.0 and mine["lr_tile"] == 32 and mine["scale"] == 16
    ref_root = "/root/reference"
    try:
        sys.path.insert(0, ref_root)
        from floodsr.preprocessing import resolve_preprocess_config as ref_resolve  # the reference itself, when present
    except Exception:
        return
    finally:
        sys.path.remove(ref_root)
    assert mine == ref_resolve(model, dem_pct_clip=99.0)
    assert resolve_preprocess_config(model) == ref_resolve(model)


def test_reconcile_tiling_rules_and_errors():
    assert reconcile_tiling(CONTRACT, {"scale": None, "lr_tile": None}, None, None) == (16, 32, 512, 8)
    assert reconcile_tiling(CONTRACT, {"scale": 8, "lr_tile": 64}, 32, 4) == (16, 32, 512, 4)  # the contract wins
    with pytest.raises(AssertionError, match="tile_size override 64 does not match model LR tile 32"):
        reconcile_tiling(CONTRACT, {}, 64, None)
    with pytest.raises(AssertionError, match="tile_overlap must be >= 0; got -1"):
        reconcile_tiling(CONTRACT, {}, None, -1)


def test_postprocess_depth_clip_and_mask():
    x = np.array([[-1.0, 0.0004, 0.001, 0.5, 7.0]], np.float32)
    assert postprocess_depth(x, 5.0, 1e-3).tolist() == [[0.0, 0.0, np.float32(0.001), 0.5, 5.0]]


def _assert_close_except_mask_flips(got, want, want_unmasked, tol=1e-4):
    """<= `tol` m wherever the 1 mm mask (ResUNet_16x_DEM.py:575-583) agrees; a pixel may only change its masked / unmasked
    state if the oracle's unmasked value lies within `tol` of the 1 mm threshold, and then it differs by < 1 mm + tol."""
    flips = (got == 0) != (want == 0)
    assert np.abs(got - want)[~flips].max() <= tol
    if flips.any():
        assert np.abs(want_unmasked[flips] - 1e-3).max() <= tol
        assert np.abs(got - want)[flips].max() <= 1e-3 + tol
    assert flips.mean() < 1e-4


@pytest.mark.gpu
def test_worker_run_prepared_matches_oracle(h1_model_fp):
    from floodsr_b200.worker import ModelWorkerB200
    from oracle.engine_ref import OracleEngine
    from oracle.stitch_np import run_tiled

    depth, dem = synth_raster(976, 1104, seed=5)
    want, n_tiles, summary = run_tiled(OracleEngine(h1_model_fp), depth, dem, window_method="feather", overlap_lr=8)
    want_unmasked = np.clip(want, 0.0, 5.0)
    want = np.where(want_unmasked < 1e-3, 0.0, want_unmasked).astype(np.float32)
    with ModelWorkerB200(h1_model_fp, precision="fp32") as worker:
        res = worker.run_prepared(depth, dem)
        assert worker.engine is not None
    assert worker.engine is None
    pre = res["preprocess"]
    assert (pre["tile_cache_size"], pre["tile_dem_stats"], pre["tile_overlap_lr"], pre["tile_size_hr"]) == (n_tiles, summary, 8, 512)
    assert pre["input_shape"]["output_shape"] == [976, 1104]
    got = res["prediction_m"]
    assert got.dtype == np.float32 and got.shape == want.shape
    _assert_close_except_mask_flips(got, want, want_unmasked)


@pytest.mark.gpu
def test_worker_run_raw_grids_matches_oracle_composition(h1_model_fp):
    """Section 8f#2: DEM on a 15x grid -> model grid (16x) -> tile loop -> prediction back on the DEM's own grid."""
    from floodsr_b200.resample import bounds_to_transform
    from floodsr_b200.worker import ModelWorkerB200
    from oracle.engine_ref import OracleEngine
    from oracle.resample_np import resample_bilinear
    from oracle.stitch_np import run_tiled

    depth, dem_model_like = synth_raster(1024, 1024, seed=11)     # depth 64 x 64
    bounds = (500.0, 800.0, 500.0 + 1024.0, 800.0 + 1024.0)
    t_raw = bounds_to_transform(*bounds, 960, 960)
    t_model = bounds_to_transform(*bounds, 1024, 1024)
    dem_raw = resample_bilinear(dem_model_like, t_model, (960, 960), t_raw)   # any terrain on the raw grid will do
    dem_raw[100:110, 200:230] = -9999.0
    # oracle composition
    dem_model = resample_bilinear(dem_raw, t_raw, (1024, 1024), t_model, -9999.0, -9999.0)
    dem_model = np.where(np.isclose(dem_model, -9999.0), 0.0, dem_model).astype(np.float32)
    pred_model, n_tiles, _ = run_tiled(OracleEngine(h1_model_fp), depth, dem_model, window_method="feather", overlap_lr=8)
    want = resample_bilinear(pred_model, t_model, (960, 960), t_raw)
    want_unmasked = np.clip(want, 0.0, 5.0)
    want = np.where(want_unmasked < 1e-3, 0.0, want_unmasked).astype(np.float32)
    with ModelWorkerB200(h1_model_fp, precision="fp32") as worker:
        res = worker.run_raw_grids(depth, bounds, dem_raw, t_raw, dem_nodata=-9999.0)
    pre = res["preprocess"]
    assert pre["resampled"] and pre["post_resampled"] and pre["tile_cache_size"] == n_tiles
    assert pre["input_shape"]["output_shape"] == [960, 960] and pre["input_shape"]["model_space_crop_height"] == 1024
    got = res["prediction_m"]
    assert got.dtype == np.float32 and got.shape == (960, 960)
    _assert_close_except_mask_flips(got, want, want_unmasked)
