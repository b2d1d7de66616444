"""Oracle restatement of the reference engine (TEST INFRASTRUCTURE ONLY).

Follows `/root/reference/floodsr/engine/ort.py:15-208` (`ModelIOContract`, `EngineORT`), with the
onnxruntime session replaced by `oracle.onnx_ref.RefSession`.
"""

from __future__ import annotations

import time
from pathlib import Path

import numpy as np
import torch

from oracle import preprocessing_np as pp
from oracle.onnx_ref import RefSession


class _OrtSession:
    """The reference's own session (`ort.InferenceSession(model, providers=[CPUExecutionProvider])`, ort.py:54,193) behind
    the three calls `OracleEngine` makes.  Only constructed where onnxruntime is importable (it is not in the offline image)."""

    def __init__(self, path, threads: int | None = None):
        import onnxruntime as ort

        so = ort.SessionOptions()
        if threads is not None:
            so.intra_op_num_threads = int(threads)
        self.sess = ort.InferenceSession(str(path), sess_options=so, providers=["CPUExecutionProvider"])
        self.version = ort.__version__

    def input_signature(self):
        return [(i.name, list(i.shape)) for i in self.sess.get_inputs()]

    def output_signature(self):
        return [(o.name, list(o.shape)) for o in self.sess.get_outputs()]

    def run(self, feeds, want=None):
        return self.sess.run(want, {k: np.asarray(v, np.float32) for k, v in feeds.items()})


def live_ort_available() -> bool:
    try:
        import onnxruntime  # noqa: F401
    except Exception:
        return False
    return True


def _hwc(dims, name):
    """ort.py:66-73."""
    assert len(dims) == 4, f"{name} must be rank-4 NHWC; got {dims}"
    h, w, c = dims[1], dims[2], dims[3]
    assert isinstance(h, int) and h > 0, f"{name} height must be fixed int; got {h}"
    assert isinstance(w, int) and w > 0, f"{name} width must be fixed int; got {w}"
    assert isinstance(c, int) and c == 1, f"{name} channels must be 1; got {c}"
    return (h, w, c)


class OracleEngine:
    """CPU oracle with the `EngineORT` surface: `.contract`-like attributes and `run_tile`."""

    def __init__(self, model_fp, dtype: torch.dtype = torch.float32, threads: int | None = None, backend: str = "auto"):
        """`backend`: "torch" = the restated interpreter (oracle.onnx_ref), "ort" = a live onnxruntime session (the reference's
        own arithmetic), "auto" = ort when importable and fp32 is asked for, else torch.  `self.backend` says which one runs."""
        self._model_fp = Path(model_fp).expanduser().resolve()
        assert self._model_fp.exists(), f"model file does not exist: {self._model_fp}"
        assert backend in ("auto", "torch", "ort")
        use_ort = backend == "ort" or (backend == "auto" and dtype == torch.float32 and live_ort_available())
        if use_ort:
            self.session = _OrtSession(self._model_fp, threads=threads)
            self.backend = f"onnxruntime {self.session.version} CPUExecutionProvider"
        else:
            self.session = RefSession(self._model_fp, dtype=dtype, threads=threads)
            self.backend = "torch-CPU restatement of the ONNX graph (onnxruntime not importable)"
        # ort.py:75-102
        ins = dict(self.session.input_signature())
        outs = self.session.output_signature()
        assert "depth_lr" in ins, "model input 'depth_lr' not found"
        assert "dem_hr" in ins, "model input 'dem_hr' not found"
        assert len(outs) > 0, "model outputs are empty"
        self.output_name = outs[0][0]
        self.depth_lr_hwc = _hwc(ins["depth_lr"], "depth_lr")
        self.dem_hr_hwc = _hwc(ins["dem_hr"], "dem_hr")
        self.output_hwc = _hwc(outs[0][1], self.output_name)
        assert self.dem_hr_hwc == self.output_hwc
        assert self.dem_hr_hwc[0] % self.depth_lr_hwc[0] == 0
        self.scale = int(self.dem_hr_hwc[0] // self.depth_lr_hwc[0])

    def model_path(self) -> Path:
        return self._model_fp

    def forward_norm(self, depth_norm: np.ndarray, dem_norm: np.ndarray) -> np.ndarray:
        """Network only: [B,32,32] + [B,512,512] normalised -> [B,512,512] normalised prediction."""
        d = np.asarray(depth_norm, dtype=np.float32)[..., None]
        e = np.asarray(dem_norm, dtype=np.float32)[..., None]
        out = self.session.run({"depth_lr": d, "dem_hr": e}, [self.output_name])[0]
        return np.asarray(out[..., 0])

    def run_tile(
        self,
        depth_lr_m,
        dem_hr_m,
        max_depth: float = 5.0,
        dem_pct_clip: float = 95.0,
        dem_ref_stats=None,
        depth_lr_nodata=None,
        dem_hr_nodata=None,
        normalize_inputs: bool = True,
        logger=None,
    ):
        """ort.py:128-208."""
        start = time.perf_counter()
        depth = np.asarray(depth_lr_m)
        dem = np.asarray(dem_hr_m)
        if normalize_inputs:
            depth = pp.replace_nodata_with_zero(depth, depth_lr_nodata)
            dem = pp.replace_nodata_with_zero(dem, dem_hr_nodata)
            assert np.isfinite(depth).all(), "low-res depth contains non-finite values after nodata replacement"
            assert np.isfinite(dem).all(), "DEM contains non-finite values after nodata replacement"
            depth_n = pp.scale_depth_log1p(depth, max_depth=float(max_depth))
            dem_n, stats = pp.normalize_dem(dem, pct_clip=float(dem_pct_clip), ref_stats=dem_ref_stats)
        else:
            depth_n = depth.astype(np.float32, copy=False)
            dem_n = dem.astype(np.float32, copy=False)
            assert np.isfinite(depth_n).all(), "low-res depth contains non-finite values"
            assert np.isfinite(dem_n).all(), "DEM contains non-finite values"
            assert float(np.min(depth_n)) >= 0.0 and float(np.max(depth_n)) <= 1.0, "depth tile must be normalized to [0, 1]"
            assert float(np.min(dem_n)) >= 0.0 and float(np.max(dem_n)) <= 1.0, "DEM tile must be normalized to [0, 1]"
            if dem_ref_stats is not None and isinstance(dem_ref_stats, dict):
                stats = {k: float(v) for k, v in dem_ref_stats.items() if k in {"p_clip", "dem_min", "dem_max"}}
            else:
                stats = {"p_clip": float(dem_pct_clip), "dem_min": 0.0, "dem_max": 1.0}
        d4 = depth_n[np.newaxis, :, :, np.newaxis].astype(np.float32, copy=False)
        e4 = dem_n[np.newaxis, :, :, np.newaxis].astype(np.float32, copy=False)
        assert d4.shape[1:] == self.depth_lr_hwc, f"depth tensor shape {d4.shape[1:]} != expected {self.depth_lr_hwc}"
        assert e4.shape[1:] == self.dem_hr_hwc, f"DEM tensor shape {e4.shape[1:]} != expected {self.dem_hr_hwc}"
        out = self.session.run({"depth_lr": d4, "dem_hr": e4}, [self.output_name])[0]
        pred_norm = np.asarray(out[0, :, :, 0], dtype=np.float32)
        pred_m = pp.invert_depth_log1p(pred_norm, max_depth=float(max_depth))
        assert pred_m.shape == self.output_hwc[:2]
        return {
            "prediction_m": pred_m.astype(np.float32, copy=False),
            "prediction_norm": pred_norm,
            "dem_stats_used": stats,
            "runtime_s": float(time.perf_counter() - start),
        }
