"""B200-aware model worker for `ResUNet_16x_DEM`.

Counterpart of the reference's `ModelWorker` (`floodsr/models/ResUNet_16x_DEM.py:108-640`) for everything that is
array work: the engine life cycle (`:128-138`), the reconciliation of scale / tile size / overlap with the model contract
(`:466-512`), the tile loop + mosaic (`:140-393`, here ONE `EngineB200.run_raster` call instead of a Python loop over
`engine.run_tile`), the final clip and low-depth mask (`:575-583`) and the `preprocess` block of the diagnostics
dictionary (`:606-640`).  Raster alignment, resampling and GeoTIFF I/O (`floodsr/preprocessing.py:285-473`, GDAL) stay on
the host in the reference code: `run_prepared` takes the prepared model-space arrays that `write_prepared_rasters`
produces, `_run_tiled_model_on_prepared` keeps the reference method's signature so that a subclass of the reference worker
can delegate to it.
"""

from __future__ import annotations

import json
import logging
import re
import time
from pathlib import Path
from typing import Any

import numpy as np

from floodsr_b200.engine.b200 import EngineB200


def resolve_preprocess_config(model_fp: str | Path, max_depth: float | None = None, dem_pct_clip: float | None = None,
                              logger=None) -> dict[str, object]:
    """Defaults from `train_config.json` beside the model (`floodsr/preprocessing.py:175-244`): same keys, same precedence."""
    log = logger or logging.getLogger(__name__)
    model_path = Path(model_fp).expanduser().resolve()
    assert model_path.exists(), f"model file does not exist: {model_path}"
    cfg: dict[str, object] = {
        "max_depth": 5.0 if max_depth is None else float(max_depth),
        "dem_pct_clip": 95.0 if dem_pct_clip is None else float(dem_pct_clip),
        "dem_ref_stats": None, "lr_tile": None, "scale": None, "model_dem_resolution": None,
    }
    train_fp = model_path.parent / "train_config.json"
    if train_fp.exists():
        train = json.loads(train_fp.read_text(encoding="utf-8"))
        if max_depth is None and train.get("max_depth") is not None:
            cfg["max_depth"] = float(train["max_depth"])
        if dem_pct_clip is None and train.get("dem_pct_clip") is not None:
            cfg["dem_pct_clip"] = float(train["dem_pct_clip"])
        stats = train.get("dem_stats") or {}
        if {"p_clip", "dem_min", "dem_max"}.issubset(stats):
            cfg["dem_ref_stats"] = {k: float(stats[k]) for k in ("dem_max", "dem_min", "p_clip")}
        shape = train.get("input_shape")
        if isinstance(shape, (tuple, list)) and len(shape) >= 2 and isinstance(shape[0], (int, float)) and float(shape[0]).is_integer():
            cfg["lr_tile"] = int(shape[0])
        if train.get("upscale") is not None:
            cfg["scale"] = int(train["upscale"])
        if train.get("dem_fp"):
            m = re.search(r"(?:^|[_/])([0-9]{2,})_?dem", str(train["dem_fp"]))
            if m is not None:
                cfg["model_dem_resolution"] = float(int(m.group(1)))
    else:
        log.debug(f"train config not found for model\n    {model_path}")
    if cfg["model_dem_resolution"] is None:
        cfg["model_dem_resolution"] = 2.0
    return cfg


def reconcile_tiling(contract, preprocess_cfg: dict, tile_size: int | None, tile_overlap: int | None, logger=None) -> tuple[int, int, int, int]:
    """`(model_scale, model_lr_tile, contract_hr_tile, overlap_lr)` with the reference's rules and errors (`:472-512`)."""
    log = logger or logging.getLogger(__name__)
    contract_scale = int(contract.scale)
    contract_lr_tile = int(contract.depth_lr_hwc[0])
    contract_hr_tile = int(contract.dem_hr_hwc[0])
    model_scale = int(preprocess_cfg["scale"]) if isinstance(preprocess_cfg.get("scale"), (int, float)) else contract_scale
    if model_scale != contract_scale:
        log.warning(f"using contract scale {contract_scale} over configured scale {model_scale}")
        model_scale = contract_scale
    model_lr_tile = int(preprocess_cfg["lr_tile"]) if isinstance(preprocess_cfg.get("lr_tile"), (int, float)) else contract_lr_tile
    if model_lr_tile != contract_lr_tile:
        log.warning(f"model config LR tile {model_lr_tile} overrides contract tile {contract_lr_tile}; "
                    "using contract tile for strict model shape checks.")
        model_lr_tile = contract_lr_tile
    if tile_size is not None:
        tile_size = int(tile_size)
        if tile_size != contract_lr_tile:
            raise AssertionError(f"tile_size override {tile_size} does not match model LR tile {contract_lr_tile}")
        model_lr_tile = tile_size
    if model_lr_tile * model_scale != contract_hr_tile:
        raise AssertionError(f"model tile mismatch: LR tile {model_lr_tile} x scale {model_scale} != contract HR tile {contract_hr_tile}")
    overlap_lr = int(tile_overlap) if tile_overlap is not None else contract_lr_tile // 4
    if overlap_lr < 0:
        raise AssertionError(f"tile_overlap must be >= 0; got {overlap_lr}")
    return model_scale, model_lr_tile, contract_hr_tile, overlap_lr


def postprocess_depth(prediction_m: np.ndarray, max_depth: float, low_depth_mask_m: float) -> np.ndarray:
    """Final clip and low-depth mask of `ModelWorker.run` (`:575-583`)."""
    out = np.clip(prediction_m, 0.0, float(max_depth)).astype(np.float32, copy=False)
    return np.where(out < float(low_depth_mask_m), 0.0, out).astype(np.float32, copy=False)


class ModelWorkerB200:
    """Worker for version `ResUNet_16x_DEM` that hands whole rasters to the B200 engine."""

    model_version = "ResUNet_16x_DEM"
    low_depth_mask_m = 1e-3

    def __init__(self, model_fp: str | Path, *, providers: tuple[str, ...] = ("B200ExecutionProvider",), logger=None,
                 precision: str | None = None):
        self.model_fp = Path(model_fp).expanduser().resolve()
        assert self.model_fp.exists(), f"model file does not exist: {self.model_fp}"
        assert providers, "providers cannot be empty"
        self.providers = tuple(providers)
        self.precision = precision
        self.log = logger or logging.getLogger(__name__)
        self.engine: EngineB200 | None = None

    def __enter__(self):
        self.engine = EngineB200(self.model_fp, providers=self.providers, logger=self.log, precision=self.precision)
        return self

    def __exit__(self, exc_type, exc, tb):
        if self.engine is not None:
            self.engine.close()
        self.engine = None
        return False

    def run_tiled_arrays(self, depth_lr_raw: np.ndarray, dem_hr_raw: np.ndarray, *, preprocess_cfg: dict, model_lr_tile: int,
                         model_scale: int, contract_hr_tile: int, window_method: str, overlap_lr: int):
        """`_run_tiled_model_on_prepared` on arrays: `(prediction_depth_m, n_tiles, tile_dem_stats_summary)`."""
        assert self.engine is not None, "worker must be entered before running inference"
        assert window_method in {"hard", "feather"}, f"unsupported window_method={window_method}"
        assert model_lr_tile * model_scale == contract_hr_tile
        return self.engine.run_raster(
            depth_lr_raw, dem_hr_raw, max_depth=float(preprocess_cfg["max_depth"]), dem_pct_clip=float(preprocess_cfg["dem_pct_clip"]),
            window_method=window_method, overlap_lr=overlap_lr,
        )

    def _run_tiled_model_on_prepared(self, *, depth_lr_fp, dem_hr_fp, preprocess_cfg, model_lr_tile, model_scale, contract_hr_tile,
                                     window_method, overlap_lr):
        """Same signature as the reference method (`ResUNet_16x_DEM.py:140-150`): reads the prepared rasters with the
        reference's own reader (needs rasterio), then one engine call."""
        from floodsr.preprocessing import _read_single_band_raster  # noqa: PLC0415 - only available next to the reference

        depth_lr_raw, _, _ = _read_single_band_raster(Path(depth_lr_fp))
        dem_hr_raw, _, _ = _read_single_band_raster(Path(dem_hr_fp))
        return self.run_tiled_arrays(depth_lr_raw, dem_hr_raw, preprocess_cfg=preprocess_cfg, model_lr_tile=model_lr_tile,
                                     model_scale=model_scale, contract_hr_tile=contract_hr_tile, window_method=window_method,
                                     overlap_lr=overlap_lr)

    def run_raw_grids(self, depth_lr: np.ndarray, depth_bounds, dem_crop: np.ndarray, dem_crop_transform, *, dem_nodata=None,
                      max_depth: float | None = None, dem_pct_clip: float | None = None, window_method: str = "feather",
                      tile_overlap: int | None = None, tile_size: int | None = None) -> dict[str, Any]:
        """`ModelWorker.run` between the raster reads and the raster write, with both grid changes on the GPU (section 8f#2).

        `depth_lr`: low-resolution depth on its native grid (nodata already replaced, `preprocessing.py:343-347`);
        `depth_bounds` = (west, south, east, north) of that raster; `dem_crop` / `dem_crop_transform`: the DEM clipped to
        those bounds on its own grid (`:352-360`).  Steps: DEM -> model grid (`:367-398`), tile loop + mosaic, prediction
        -> raw DEM grid (`ResUNet_16x_DEM.py:552-573`), clip and low-depth mask (`:575-583`).
        """
        from floodsr_b200.resample import align_dem_to_model_grid, prediction_to_raw_grid  # noqa: PLC0415

        start = time.perf_counter()
        assert self.engine is not None, "worker must be used under context management"
        depth_lr = np.asarray(depth_lr, dtype=np.float32)
        dem_crop = np.asarray(dem_crop, dtype=np.float32)
        assert dem_crop.size > 0, "clipped DEM is empty"
        if not np.isfinite(depth_lr).all():
            raise AssertionError("low-res depth contains non-finite values")
        if depth_lr.min() < 0.0:
            raise AssertionError(f"low-res depth has negative values: min={float(depth_lr.min())}")
        if self.engine.contract is None:
            self.engine.load()
        scale = int(self.engine.contract.scale)
        aligned = align_dem_to_model_grid(self.engine, dem_crop, dem_crop_transform, depth_bounds, depth_lr.shape, scale, dem_nodata)
        res = self.run_prepared(depth_lr, aligned["dem_hr"], max_depth=max_depth, dem_pct_clip=dem_pct_clip, window_method=window_method,
                                tile_overlap=tile_overlap, tile_size=tile_size, _postprocess=False)
        prediction_model_m = res["prediction_m"]
        post_resampled = tuple(dem_crop.shape) != tuple(prediction_model_m.shape)
        prediction_out_m = prediction_to_raw_grid(self.engine, prediction_model_m, aligned["dem_hr_transform"], dem_crop.shape, dem_crop_transform)
        res["prediction_m"] = postprocess_depth(prediction_out_m, float(res["preprocess"]["max_depth"]), self.low_depth_mask_m)
        res["runtime_s"] = float(time.perf_counter() - start)
        res["preprocess"]["resampled"] = bool(aligned["resampled"])
        res["preprocess"]["post_resampled"] = bool(post_resampled)
        res["preprocess"]["input_shape"].update(
            crop_height=int(res["prediction_m"].shape[0]), crop_width=int(res["prediction_m"].shape[1]),
            output_shape=[int(x) for x in res["prediction_m"].shape],
        )
        return res

    def run_prepared(self, depth_lr_prepared: np.ndarray, dem_hr_prepared: np.ndarray, *, max_depth: float | None = None,
                     dem_pct_clip: float | None = None, window_method: str = "feather", tile_overlap: int | None = None,
                     tile_size: int | None = None, _postprocess: bool = True) -> dict[str, Any]:
        """The array part of `ModelWorker.run` for prepared (aligned, model-space) rasters."""
        start = time.perf_counter()
        assert self.engine is not None, "worker must be used under context management"
        window_method = (window_method or "feather").strip().lower()
        assert window_method in {"hard", "feather"}, f"unsupported window_method={window_method}"
        cfg = resolve_preprocess_config(self.model_fp, max_depth=max_depth, dem_pct_clip=dem_pct_clip, logger=self.log)
        assert self.engine.contract is not None, "engine contract must be available"
        model_scale, model_lr_tile, contract_hr_tile, overlap_lr = reconcile_tiling(self.engine.contract, cfg, tile_size, tile_overlap, self.log)
        prediction_model_m, n_tiles, tile_dem_stats = self.run_tiled_arrays(
            depth_lr_prepared, dem_hr_prepared, preprocess_cfg=cfg, model_lr_tile=model_lr_tile, model_scale=model_scale,
            contract_hr_tile=contract_hr_tile, window_method=window_method, overlap_lr=overlap_lr,
        )
        dem_shape = tuple(np.asarray(dem_hr_prepared).shape)
        assert prediction_model_m.shape == dem_shape, f"prediction shape {prediction_model_m.shape} must match preprocessed DEM shape {dem_shape}"
        prediction_out_m = postprocess_depth(prediction_model_m, float(cfg["max_depth"]), self.low_depth_mask_m) if _postprocess else prediction_model_m
        return {
            "prediction_m": prediction_out_m,
            "runtime_s": float(time.perf_counter() - start),
            "model_version": self.model_version,
            "model_fp": str(self.model_fp),
            "preprocess": {
                "max_depth": float(cfg["max_depth"]),
                "dem_pct_clip": float(cfg["dem_pct_clip"]),
                "dem_ref_stats": cfg["dem_ref_stats"],
                "window_method": window_method,
                "tile_overlap_lr": overlap_lr,
                "tile_size_lr": model_lr_tile,
                "tile_size_hr": contract_hr_tile,
                "model_scale": model_scale,
                "tile_cache_size": n_tiles,
                "tile_dem_stats": tile_dem_stats,
                "input_shape": {
                    "crop_height": int(prediction_out_m.shape[0]),
                    "crop_width": int(prediction_out_m.shape[1]),
                    "model_space_crop_height": int(prediction_model_m.shape[0]),
                    "model_space_crop_width": int(prediction_model_m.shape[1]),
                    "aligned_depth_shape": [int(x) for x in np.asarray(depth_lr_prepared).shape],
                    "aligned_dem_shape": [int(x) for x in dem_shape],
                    "output_shape": [int(x) for x in dem_shape],
                },
            },
        }
