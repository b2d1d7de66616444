"""`EngineB200`: the B200 engine backend behind floodsr's engine contract.

Drop-in for the reference's `EngineORT` (`floodsr/engine/ort.py:28-208`): same constructor signature,
attributes (`contract`, `providers`, `session`), `load/close/model_path/run_tile` methods, result keys,
AssertionError conditions and messages.  The arithmetic runs in `libfloodsr_b200.so` (hand-written
sm_100a CUDA) through the C ABI of `include/floodsr_b200.h`; this module only validates arguments, resolves
host-side scalars and moves numpy arrays across ctypes.

Beyond the reference contract it adds the batched and whole-raster entry points the throughput path
needs (`run_tiles`, `run_raster`), which replace the Python tile loop of
`ModelWorker._run_tiled_model_on_prepared` (`floodsr/models/ResUNet_16x_DEM.py:140-393`) by one call.
"""

from __future__ import annotations

import ctypes as C
import logging
import os
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from floodsr_b200 import _lib
from floodsr_b200.engine.base import EngineBase
from floodsr_b200.graph import LoweredModel, lower_onnx
from floodsr_b200.preprocessing import check_pct_clip, depth_log1p_denom, nodata_tolerance, percentile_ranks
from floodsr_b200.tiling import build_feather_ramp, window_grid


@dataclass(frozen=True)
class ModelIOContract:
    """Resolved model tensor names and spatial dimensions (same fields as `ort.py:15-25`)."""

    depth_input_name: str
    dem_input_name: str
    output_name: str
    depth_lr_hwc: tuple[int, int, int]
    dem_hr_hwc: tuple[int, int, int]
    output_hwc: tuple[int, int, int]
    scale: int


_PRECISIONS = {"fp32": _lib.PREC_FP32, "fp16": _lib.PREC_FP16, "fp32_simt": _lib.PREC_FP32_SIMT}


def _parse_ref_stats(ref_stats: dict[str, float]) -> tuple[float, float, float]:
    """`_parse_dem_normalization_stats` (`floodsr/preprocessing.py:39-58`): same checks, same wording."""
    missing = {"p_clip", "dem_min", "dem_max"}.difference(ref_stats.keys())
    if missing:
        raise AssertionError(f"DEM ref_stats missing keys: {sorted(missing)}")
    p_clip, dem_min, dem_max = float(ref_stats["p_clip"]), float(ref_stats["dem_min"]), float(ref_stats["dem_max"])
    if not (np.isfinite(p_clip) and np.isfinite(dem_min) and np.isfinite(dem_max)):
        raise AssertionError("DEM ref_stats values must be finite")
    if p_clip < 0:
        raise AssertionError(f"DEM p_clip must be >= 0; got {p_clip}")
    if dem_min > dem_max:
        raise AssertionError(f"DEM dem_min must be <= dem_max; got min={dem_min} max={dem_max}")
    if (dem_max - dem_min) <= 0:
        raise AssertionError(f"DEM range must be > 0; got min={dem_min}, max={dem_max}")
    return p_clip, dem_min, dem_max


class EngineB200(EngineBase):
    """B200 engine: ONNX graph lowered to fused sm_100a kernels, inputs normalised and stitched on the GPU."""

    def __init__(
        self,
        model_fp: str | Path,
        providers: tuple[str, ...] = ("B200ExecutionProvider",),
        logger=None,
        *,
        precision: str | None = None,
        device: int | None = None,
    ):
        """Parse the model, lower it and create the device engine.

        `providers` is accepted for signature parity with `EngineORT` (`ort.py:31-36`); an entry of the form
        `"B200ExecutionProvider:fp16"` selects the precision mode.  `precision` (or the environment variable
        FLOODSR_B200_PRECISION) overrides it.  Both modes run on tcgen05 tensor cores:
        "fp32" (default): the reference's tolerance, <= 1e-4 m against the fp32 CPU path (split fp16 operands, three MMAs
        per product); "fp16": the 16-bit mode, <= 1e-2 m and the same wet/dry mask at 0.01 m (measured 2.6e-3 m on the
        H1 graph), 2.6x the throughput.  "fp32_simt" runs plain fp32 FMA on the CUDA cores (diagnostic).  bf16 operands
        measured 2.4e-2 m, outside the 16-bit bound, and are not offered.
        """
        self._model_fp = Path(model_fp).expanduser().resolve()
        assert self._model_fp.exists(), f"model file does not exist: {self._model_fp}"
        assert providers, "providers cannot be empty"
        self.providers = tuple(providers)
        self.log = logger or logging.getLogger(__name__)
        prec = precision or os.environ.get("FLOODSR_B200_PRECISION")
        if prec is None:
            for p in self.providers:
                if isinstance(p, str) and p.startswith("B200ExecutionProvider:"):
                    prec = p.split(":", 1)[1]
        self.precision = (prec or "fp32").lower()
        if self.precision not in _PRECISIONS:
            raise ValueError(f"unknown precision '{self.precision}'; expected one of {sorted(_PRECISIONS)}")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", os.environ.get("FLOODSR_B200_DEVICE", "0")))
        self.device = int(device)
        self._handle: C.c_void_p | None = None
        self.session: Any = None
        self.contract: ModelIOContract | None = None
        self.lowered: LoweredModel | None = None
        self.load()

    # -- lifecycle -------------------------------------------------------------------------------
    def model_path(self) -> Path:
        """Return the model path used by this engine."""
        return self._model_fp

    def load(self) -> None:
        """Lower the ONNX graph and upload plan + weights (replaces `InferenceSession(...)`, `ort.py:51-59`)."""
        self.close()
        self.log.debug(f"lowering ONNX model for the B200 engine from\n    {self._model_fp}")
        lowered = lower_onnx(self._model_fp)
        lib = _lib.load_library()
        handle = C.c_void_p()
        plan = (C.c_char * len(lowered.plan_bytes)).from_buffer_copy(lowered.plan_bytes)
        weights = np.ascontiguousarray(lowered.weights, dtype=np.float32)
        assert np.isfinite(weights).all(), f"model weights contain non-finite values: {self._model_fp}"
        _lib.check(
            lib.fsr_create(
                C.cast(plan, C.c_void_p), len(lowered.plan_bytes), _lib.fptr(weights), weights.size, self.device,
                _PRECISIONS[self.precision], C.byref(handle),
            )
        )
        self._handle = handle
        self.session = self  # truthy stand-in for code that checks `engine.session is not None`
        self.lowered = lowered
        c = lowered.contract
        self.contract = ModelIOContract(
            c.depth_input_name, c.dem_input_name, c.output_name, c.depth_lr_hwc, c.dem_hr_hwc, c.output_hwc, c.scale
        )
        self.log.info(
            f"loaded B200 engine for '{self._model_fp.name}' on cuda:{self.device} precision={self.precision} "
            f"({len(lowered.ops)} fused layers, {lowered.macs_per_tile() / 1e9:.3f} GMAC/tile) and scale={c.scale}"
        )

    def close(self) -> None:
        """Release device resources (mirrors `EngineORT.close`, `ort.py:61-64`)."""
        if getattr(self, "_handle", None):
            _lib.load_library().fsr_destroy(self._handle)
        self._handle = None
        self.session = None
        self.contract = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- diagnostics -----------------------------------------------------------------------------
    def launch_count(self) -> int:
        """CUDA kernel launches issued by this engine so far."""
        return int(_lib.load_library().fsr_launch_count(self._handle)) if self._handle else 0

    def profile(self, on: bool) -> None:
        """Enable/disable (and reset) per-stage CUDA-event timing inside the engine."""
        _lib.check(_lib.load_library().fsr_profile_enable(self._handle, 1 if on else 0))

    def profile_fetch(self) -> dict[str, tuple[float, int]]:
        """{stage: (device ms, launch groups)} since `profile(True)`."""
        n = len(_lib.PROF_CATEGORIES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(_lib.load_library().fsr_profile_fetch(self._handle, ms, cnt, n))
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.PROF_CATEGORIES)}

    def debug_tensor(self, tensor: int, n_tiles: int = 1) -> np.ndarray:
        """Intermediate activation `tensor` of the last forward pass as float32 [n_tiles, h, w, c]."""
        lib = _lib.load_library()
        h, w, c = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.fsr_debug_tensor_shape(self._handle, tensor, C.byref(h), C.byref(w), C.byref(c)))
        out = np.empty((n_tiles, h.value, w.value, c.value), dtype=np.float32)
        _lib.check(lib.fsr_debug_read_tensor(self._handle, tensor, n_tiles, _lib.fptr(out)))
        return out

    def macs_per_tile(self) -> int:
        assert self.lowered is not None
        return self.lowered.macs_per_tile()

    # -- parameter resolution --------------------------------------------------------------------
    def _tile_params(
        self,
        max_depth: float,
        dem_pct_clip: float,
        dem_ref_stats,
        depth_lr_nodata,
        dem_hr_nodata,
        normalize_inputs: bool,
    ) -> _lib.TileParams:
        assert self.contract is not None, "model contract must be available before inference"
        p = _lib.TileParams()
        max_depth = float(max_depth)
        p.max_depth = max_depth
        p.depth_denom = depth_log1p_denom(max_depth)
        p.normalize_inputs = 1 if normalize_inputs else 0
        p.dem_pct_clip = float(dem_pct_clip)
        n_px = self.contract.dem_hr_hwc[0] * self.contract.dem_hr_hwc[1]
        p.rank_lo, p.rank_hi, p.gamma = 0, 0, 0.0
        if normalize_inputs:
            p.has_depth_nodata, p.depth_nodata, p.depth_nodata_tol = nodata_tolerance(depth_lr_nodata)
            p.has_dem_nodata, p.dem_nodata, p.dem_nodata_tol = nodata_tolerance(dem_hr_nodata)
            if dem_ref_stats is None:
                p.rank_lo, p.rank_hi, p.gamma = percentile_ranks(n_px, check_pct_clip(dem_pct_clip))
            else:
                p.has_ref_stats = 1
                p.ref_p_clip, p.ref_dem_min, p.ref_dem_max = _parse_ref_stats(dem_ref_stats)
        return p

    @staticmethod
    def _raise_flags(flags: int, normalize_inputs: bool, stats: np.ndarray | None) -> None:
        """Re-raise device-side validation failures with the reference's AssertionError messages."""
        if not flags:
            return
        suffix = " after nodata replacement" if normalize_inputs else ""
        if flags & _lib.FLAG_DEPTH_NONFINITE:
            raise AssertionError(f"low-res depth contains non-finite values{suffix}")
        if flags & _lib.FLAG_DEM_NONFINITE:
            raise AssertionError(f"DEM contains non-finite values{suffix}")
        if flags & _lib.FLAG_DEPTH_NOT_UNIT:
            raise AssertionError("depth tile must be normalized to [0, 1]")
        if flags & _lib.FLAG_DEM_NOT_UNIT:
            raise AssertionError("DEM tile must be normalized to [0, 1]")
        if flags & _lib.FLAG_DEM_FLAT_NONZERO:
            bad = None
            if stats is not None:
                s = np.asarray(stats, dtype=np.float64).reshape(-1, 3)
                idx = np.nonzero((s[:, 2] - s[:, 1] <= 0) & ~(np.isclose(s[:, 1], 0.0)))[0]
                if idx.size:
                    bad = s[idx[0]]
            if bad is not None:
                raise AssertionError(f"DEM range must be > 0; got min={float(bad[1])}, max={float(bad[2])}")
            raise AssertionError("DEM range must be > 0")
        if flags & _lib.FLAG_PRED_NONFINITE:
            raise FloatingPointError(
                "the network produced non-finite values (a 16-bit activation overflowed, or the weights are not finite); "
                "the reference's clip would hide them as 0 m -- use precision='fp32' or check the model file")
        raise AssertionError(f"device-side input validation failed (flags={flags:#x})")

    def _call(self, code: int, flags: C.c_uint32, normalize_inputs: bool, stats: np.ndarray | None) -> None:
        if code == _lib.FSR_E_ASSERT:
            self._raise_flags(int(flags.value), normalize_inputs, stats)
        _lib.check(code)

    # -- EngineBase.run_tile ---------------------------------------------------------------------
    def run_tile(
        self,
        depth_lr_m: np.ndarray,
        dem_hr_m: np.ndarray,
        max_depth: float = 5.0,
        dem_pct_clip: float = 95.0,
        dem_ref_stats: dict[str, float] | None = None,
        depth_lr_nodata: float | None = None,
        dem_hr_nodata: float | None = None,
        normalize_inputs: bool = True,
        logger=None,
    ) -> dict[str, Any]:
        """One tile from prepared depth/DEM arrays to predicted depth (same contract as `ort.py:128-208`)."""
        assert self._handle, "session must be loaded before inference"
        assert self.contract is not None, "model contract must be available before inference"
        start = time.perf_counter()
        depth = np.ascontiguousarray(np.asarray(depth_lr_m), dtype=np.float32)
        dem = np.ascontiguousarray(np.asarray(dem_hr_m), dtype=np.float32)
        assert depth.shape + (1,) == self.contract.depth_lr_hwc, (
            f"depth tensor shape {depth.shape + (1,)} != expected {self.contract.depth_lr_hwc}"
        )
        assert dem.shape + (1,) == self.contract.dem_hr_hwc, (
            f"DEM tensor shape {dem.shape + (1,)} != expected {self.contract.dem_hr_hwc}"
        )
        res = self.run_tiles(
            depth[None], dem[None], max_depth=max_depth, dem_pct_clip=dem_pct_clip, dem_ref_stats=dem_ref_stats,
            depth_lr_nodata=depth_lr_nodata, dem_hr_nodata=dem_hr_nodata, normalize_inputs=normalize_inputs,
        )
        pred_m = res["prediction_m"][0]
        assert pred_m.shape == self.contract.output_hwc[:2], (
            f"prediction shape {pred_m.shape} != expected {self.contract.output_hwc[:2]}"
        )
        return {
            "prediction_m": pred_m,
            "prediction_norm": res["prediction_norm"][0],
            "dem_stats_used": res["dem_stats_used"][0],
            "runtime_s": float(time.perf_counter() - start),
        }

    # -- batched tiles (extension) ---------------------------------------------------------------
    def run_tiles(
        self,
        depth_lr_m: np.ndarray,
        dem_hr_m: np.ndarray,
        max_depth: float = 5.0,
        dem_pct_clip: float = 95.0,
        dem_ref_stats: dict[str, float] | None = None,
        depth_lr_nodata: float | None = None,
        dem_hr_nodata: float | None = None,
        normalize_inputs: bool = True,
        want_norm: bool = True,
    ) -> dict[str, Any]:
        """`run_tile` for a batch: depth [B, lr, lr], DEM [B, hr, hr] -> dict of batched arrays."""
        assert self._handle, "session must be loaded before inference"
        assert self.contract is not None, "model contract must be available before inference"
        start = time.perf_counter()
        depth = np.ascontiguousarray(np.asarray(depth_lr_m), dtype=np.float32)
        dem = np.ascontiguousarray(np.asarray(dem_hr_m), dtype=np.float32)
        lr, hr = self.contract.depth_lr_hwc[0], self.contract.dem_hr_hwc[0]
        assert depth.ndim == 3 and depth.shape[1:] == (lr, lr), f"depth batch shape {depth.shape} != (B, {lr}, {lr})"
        assert dem.ndim == 3 and dem.shape[1:] == (hr, hr), f"DEM batch shape {dem.shape} != (B, {hr}, {hr})"
        assert depth.shape[0] == dem.shape[0] and depth.shape[0] > 0, "depth and DEM batches differ in length"
        b = depth.shape[0]
        params = self._tile_params(max_depth, dem_pct_clip, dem_ref_stats, depth_lr_nodata, dem_hr_nodata, normalize_inputs)
        pred_m = np.empty((b, hr, hr), dtype=np.float32)
        pred_norm = np.empty((b, hr, hr), dtype=np.float32) if want_norm else None
        stats = np.zeros((b, 3), dtype=np.float32)
        flags = C.c_uint32(0)
        code = _lib.load_library().fsr_run_tiles(
            self._handle, _lib.fptr(depth), _lib.fptr(dem), b, C.byref(params), _lib.fptr(pred_m),
            _lib.fptr(pred_norm) if pred_norm is not None else None, _lib.fptr(stats), C.byref(flags),
        )
        self._call(code, flags, normalize_inputs, stats)
        if not normalize_inputs:
            # ort.py:175-180
            if dem_ref_stats is not None and isinstance(dem_ref_stats, dict):
                one = {k: float(v) for k, v in dem_ref_stats.items() if k in {"p_clip", "dem_min", "dem_max"}}
            else:
                one = {"p_clip": float(dem_pct_clip), "dem_min": 0.0, "dem_max": 1.0}
            stats_l = [dict(one) for _ in range(b)]
        else:
            stats_l = [{"p_clip": float(s[0]), "dem_min": float(s[1]), "dem_max": float(s[2])} for s in stats]
        return {
            "prediction_m": pred_m,
            "prediction_norm": pred_norm,
            "dem_stats_used": stats_l,
            "runtime_s": float(time.perf_counter() - start),
        }

    # -- whole raster (extension; replaces the Python tile loop) ---------------------------------
    def run_raster(
        self,
        depth_lr_raw: np.ndarray,
        dem_hr_raw: np.ndarray,
        *,
        max_depth: float = 5.0,
        dem_pct_clip: float = 95.0,
        window_method: str = "feather",
        overlap_lr: int | None = None,
        out: np.ndarray | None = None,
    ) -> tuple[np.ndarray, int, dict[str, float] | None]:
        """Tile, normalise, predict and stitch a prepared model-space raster pair in one device pass.

        Same inputs/outputs as `_run_tiled_model_on_prepared` (`ResUNet_16x_DEM.py:140-393`) with arrays in
        place of GeoTIFF paths: returns `(prediction_depth_m [H, W] float32, n_tiles, tile_dem_stats_summary)`.
        """
        assert self._handle, "worker must be entered before running inference"
        assert self.contract is not None, "engine contract must be available"
        assert window_method in {"hard", "feather"}, f"unsupported window_method={window_method}"
        depth = np.ascontiguousarray(np.asarray(depth_lr_raw), dtype=np.float32)
        dem = np.ascontiguousarray(np.asarray(dem_hr_raw), dtype=np.float32)
        assert depth.ndim == 2, f"aligned depth must be 2D; got {depth.shape}"
        assert dem.ndim == 2, f"aligned DEM must be 2D; got {dem.shape}"
        scale = self.contract.scale
        lr_tile, hr_tile = self.contract.depth_lr_hwc[0], self.contract.dem_hr_hwc[0]
        crop_h, crop_w = dem.shape
        exp_h, exp_w = crop_h // scale, crop_w // scale
        assert exp_h > 0 and exp_w > 0, (
            f"expected low-resolution shape invalid {(exp_h, exp_w)} from crop {(crop_h, crop_w)} and scale={scale}"
        )
        assert depth.shape == (exp_h, exp_w), (
            f"depth shape {depth.shape} does not match crop/scale target {(exp_h, exp_w)}"
        )
        if overlap_lr is None:
            overlap_lr = lr_tile // 4  # ResUNet_16x_DEM.py:510
        overlap_hr = int(overlap_lr) * scale
        ys, xs = window_grid(crop_h, crop_w, hr_tile, window_method, overlap_hr)
        ramp = build_feather_ramp(hr_tile, overlap_hr) if window_method == "feather" else None
        params = self._tile_params(max_depth, dem_pct_clip, None, None, None, True)
        ys_a, xs_a = np.asarray(ys, dtype=np.int32), np.asarray(xs, dtype=np.int32)
        if out is None:
            out = np.empty((crop_h, crop_w), dtype=np.float32)
        assert out.shape == (crop_h, crop_w) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        stats = np.zeros((len(ys) * len(xs), 3), dtype=np.float32)
        flags = C.c_uint32(0)
        code = _lib.load_library().fsr_run_raster(
            self._handle, _lib.fptr(depth), _lib.fptr(dem), crop_h, crop_w,
            _lib.WINDOW_FEATHER if window_method == "feather" else _lib.WINDOW_HARD, overlap_hr,
            _lib.iptr(ys_a), len(ys), _lib.iptr(xs_a), len(xs), _lib.fptr(ramp) if ramp is not None else None,
            C.byref(params), _lib.fptr(out), _lib.fptr(stats), C.byref(flags),
        )
        if code == _lib.FSR_E_ASSERT and int(flags.value) & (_lib.FLAG_DEPTH_NONFINITE | _lib.FLAG_DEM_NONFINITE):
            # the worker asserts finiteness of the whole prepared rasters up front (ResUNet_16x_DEM.py:188-189)
            if int(flags.value) & _lib.FLAG_DEPTH_NONFINITE:
                raise AssertionError("aligned depth contains non-finite values")
            raise AssertionError("aligned DEM contains non-finite values")
        self._call(code, flags, True, stats)
        return out, len(ys) * len(xs), summarise_tile_stats(stats)

    # -- stage-level entry points (parity tests of the individual kernels) -------------------------
    def stage_normalize(self, depth_lr_m, dem_hr_m, **kwargs) -> dict[str, Any]:
        """a5-a8 only: nodata, finite checks, depth log1p scaling, DEM stats + normalisation for [B,...] tiles."""
        depth = np.ascontiguousarray(np.asarray(depth_lr_m), dtype=np.float32)
        dem = np.ascontiguousarray(np.asarray(dem_hr_m), dtype=np.float32)
        b = dem.shape[0]
        normalize_inputs = kwargs.get("normalize_inputs", True)
        params = self._tile_params(
            kwargs.get("max_depth", 5.0), kwargs.get("dem_pct_clip", 95.0), kwargs.get("dem_ref_stats"),
            kwargs.get("depth_lr_nodata"), kwargs.get("dem_hr_nodata"), normalize_inputs,
        )
        depth_n, dem_n = np.empty_like(depth), np.empty_like(dem)
        stats = np.zeros((b, 3), dtype=np.float32)
        flags = C.c_uint32(0)
        code = _lib.load_library().fsr_stage_normalize(
            self._handle, _lib.fptr(depth), _lib.fptr(dem), b, C.byref(params), _lib.fptr(depth_n), _lib.fptr(dem_n),
            _lib.fptr(stats), C.byref(flags),
        )
        self._call(code, flags, normalize_inputs, stats)
        return {"depth_norm": depth_n, "dem_norm": dem_n, "stats": stats}

    def stage_forward(self, depth_norm, dem_norm) -> np.ndarray:
        """a10 only: network forward on normalised [B,lr,lr] + [B,hr,hr] -> normalised prediction [B,hr,hr]."""
        depth = np.ascontiguousarray(np.asarray(depth_norm), dtype=np.float32)
        dem = np.ascontiguousarray(np.asarray(dem_norm), dtype=np.float32)
        out = np.empty_like(dem)
        _lib.check(_lib.load_library().fsr_stage_forward(self._handle, _lib.fptr(depth), _lib.fptr(dem), dem.shape[0], _lib.fptr(out)))
        return out

    def stage_invert(self, pred_norm, max_depth: float = 5.0) -> np.ndarray:
        """a11 only: invert_depth_log1p_np."""
        x = np.ascontiguousarray(np.asarray(pred_norm), dtype=np.float32)
        out = np.empty_like(x)
        _lib.check(
            _lib.load_library().fsr_stage_invert(
                self._handle, _lib.fptr(x), x.size, float(max_depth), depth_log1p_denom(max_depth), _lib.fptr(out)
            )
        )
        return out

    def stage_blend(self, tiles, h: int, w: int, window_method: str = "feather", overlap_lr: int = 8, max_depth: float = 5.0,
                    y_starts=None, x_starts=None) -> np.ndarray:
        """a15 mosaic only: per-window predictions [ny*nx, hr, hr] in window order -> stitched [h, w] raster."""
        assert self.contract is not None
        hr_tile = self.contract.dem_hr_hwc[0]
        overlap_hr = int(overlap_lr) * self.contract.scale
        ys, xs = window_grid(h, w, hr_tile, window_method, overlap_hr)
        if y_starts is not None:
            ys, xs = list(y_starts), list(x_starts)
        t = np.ascontiguousarray(np.asarray(tiles), dtype=np.float32)
        assert t.shape == (len(ys) * len(xs), hr_tile, hr_tile), t.shape
        ramp = build_feather_ramp(hr_tile, overlap_hr) if window_method == "feather" else None
        ys_a, xs_a = np.asarray(ys, np.int32), np.asarray(xs, np.int32)
        out = np.empty((h, w), dtype=np.float32)
        _lib.check(
            _lib.load_library().fsr_stage_blend(
                self._handle, _lib.fptr(t), h, w, _lib.WINDOW_FEATHER if window_method == "feather" else _lib.WINDOW_HARD,
                overlap_hr, _lib.iptr(ys_a), len(ys), _lib.iptr(xs_a), len(xs), _lib.fptr(ramp) if ramp is not None else None,
                float(max_depth), _lib.fptr(out),
            )
        )
        return out


def summarise_tile_stats(stats: np.ndarray) -> dict[str, float] | None:
    """Per-raster summary of tile DEM stats, as `ResUNet_16x_DEM.py:366-389` computes it (float32)."""
    if stats is None or len(stats) == 0:
        return None
    st = np.asarray(stats, dtype=np.float32).reshape(-1, 3)
    rng = st[:, 2] - st[:, 1]
    return {
        "tile_count": float(st.shape[0]),
        "dem_p_clip_min": float(st[:, 0].min()),
        "dem_p_clip_mean": float(st[:, 0].mean()),
        "dem_p_clip_max": float(st[:, 0].max()),
        "dem_range_min": float(rng.min()),
        "dem_range_mean": float(rng.mean()),
        "dem_range_max": float(rng.max()),
    }


def install_as_floodsr_engine(precision: str | None = None) -> None:
    """Bind this backend in place of the reference's ORT engine: `floodsr.engine.EngineORT = EngineB200`.

    The reference worker re-imports `from floodsr.engine import EngineORT` every time it is resolved
    (`floodsr/model_registry.py:373-394`, `floodsr/models/ResUNet_16x_DEM.py:80`), so rebinding the name (or,
    when `floodsr.engine` cannot be imported because onnxruntime is absent, registering shim modules under
    the same names) routes `floodsr tohr` through the B200 backend with no edit to the reference.
    """
    import importlib
    import sys
    import types

    if precision is not None:
        os.environ["FLOODSR_B200_PRECISION"] = precision
    import floodsr_b200.engine as mine
    import floodsr_b200.engine.base as mine_base
    import floodsr_b200.engine.providers as mine_prov

    try:
        pkg = importlib.import_module("floodsr.engine")
    except ImportError:
        importlib.import_module("floodsr")  # the parent package must exist
        pkg = types.ModuleType("floodsr.engine")
        pkg.__path__ = []  # mark as package
        sys.modules["floodsr.engine"] = pkg
        sys.modules["floodsr.engine.base"] = mine_base
        sys.modules["floodsr.engine.providers"] = mine_prov
        ort_mod = types.ModuleType("floodsr.engine.ort")
        ort_mod.EngineORT = EngineB200
        ort_mod.ModelIOContract = ModelIOContract
        sys.modules["floodsr.engine.ort"] = ort_mod
        pkg.get_onnxruntime_info = mine.get_onnxruntime_info
        pkg.get_rasterio_info = mine.get_rasterio_info
        pkg.EngineBase = mine_base.EngineBase
    pkg.EngineORT = EngineB200
    pkg.EngineB200 = EngineB200
