"""Host-side logic of the product package (CPU): geometry mirrors, percentile/nodata scalars, C-ABI surface."""

from __future__ import annotations

import ast
import re
from pathlib import Path

import numpy as np
import pytest

from floodsr_b200 import _lib, tiling
from floodsr_b200.preprocessing import check_pct_clip, depth_log1p_denom, nodata_tolerance, percentile_ranks

REPO = Path(__file__).resolve().parents[1]


def test_tiling_mirror_matches_reference_golden(golden):
    meta, arrays = golden
    for key, want in meta["tile_starts"].items():
        total, tile, stride = (int(v) for v in key.split(","))
        assert tiling.build_tile_starts(total, tile, stride) == want, key
    for key in [k for k in arrays.files if k.startswith("ramp/")]:
        tile, ov = (int(v) for v in key[5:].split(","))
        assert tiling.build_feather_ramp(tile, ov).tobytes() == arrays[key].tobytes(), key
    assert [list(t) for t in tiling.iter_window_origins([0, 384, 512], [0, 100])] == meta["origins_3x2"]


def test_tiling_assertions_like_reference():
    with pytest.raises(AssertionError, match="total_size must be > 0"):
        tiling.build_tile_starts(0, 512, 384)
    with pytest.raises(AssertionError, match="overlap must be < tile_size"):
        tiling.build_feather_ramp(512, 512)
    with pytest.raises(AssertionError, match="feather windowing requires overlap_lr > 0"):
        tiling.window_grid(1024, 1024, 512, "feather", 0)
    # tile counts quoted in SURVEY.md section 8a for the BASELINE configs
    for size, n in [(4096, 11), (8192, 21), (32768, 85)]:
        ys, xs = tiling.window_grid(size, size, 512, "feather", 128)
        assert len(ys) == len(xs) == n
    assert tiling.window_grid(976, 1104, 512, "hard", 128) == ([0, 512], [0, 512, 1024])


def test_split_tile_rows_covers_grid():
    assert [b - a for a, b in tiling.split_tile_rows(85, 8)] == [11, 11, 11, 11, 11, 10, 10, 10]
    for n, p in [(1, 1), (3, 8), (21, 4), (85, 8)]:
        parts = tiling.split_tile_rows(n, p)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(len(parts) - 1))


def test_percentile_ranks_reproduce_numpy_float32_path():
    assert percentile_ranks(262144, 95.0) == (249035, 249036, 0.84375)
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.choice([262144, 1000, 4096, 777]))
        pct = float(rng.choice([95.0, 50.0, 100.0, 99.9, 0.5, 12.34]))
        x = (rng.random(n) * float(rng.choice([1, 1000, 1e-3])) + float(rng.choice([0, 1073.0]))).astype(np.float32)
        lo, hi, g = percentile_ranks(n, pct)
        s = np.sort(x)
        a, b, t = s[lo], s[hi], np.float32(g)
        d = np.float32(b - a)
        r = np.float32(a + np.float32(d * t))
        if t >= 0.5:
            r = np.float32(b - np.float32(d * np.float32(np.float32(1) - t)))
        assert float(r) == float(np.nanpercentile(x, pct)), (n, pct)


def test_nodata_tolerance_matches_isclose():
    rng = np.random.default_rng(1)
    for nodata in [-9999.0, 0.0, 1e-9, 3.4e38, -32768.0, 255.0]:
        has, nd32, tol = nodata_tolerance(nodata)
        x = np.concatenate([np.float32(nodata) + rng.normal(0, abs(nodata) * 2e-5 + 1e-8, 2000).astype(np.float32),
                            rng.normal(0, 100, 500).astype(np.float32)]).astype(np.float32)
        mine = (x == np.float32(nd32)) | (np.abs(x - np.float32(nd32)) <= np.float32(tol))
        assert has == 1 and np.array_equal(mine, np.isclose(x, nodata)), nodata
    assert nodata_tolerance(None) == (0, 0.0, -1.0)
    assert nodata_tolerance(float("nan"))[2] == -1.0


def test_scalar_validation_wording():
    assert depth_log1p_denom(5.0) == float(np.log1p(5.0))
    with pytest.raises(AssertionError, match="max_depth must be finite and > 0"):
        depth_log1p_denom(0.0)
    with pytest.raises(AssertionError, match=r"dem_pct_clip must be finite and in \(0, 100\]"):
        check_pct_clip(0.0)
    with pytest.raises(AssertionError):
        check_pct_clip(100.5)


def test_c_abi_exports_every_declared_symbol():
    """The in-tree library loads here (no GPU needed) and exports exactly what include/floodsr_b200.h declares."""
    from floodsr_b200.build import build

    build()
    header = (REPO / "include" / "floodsr_b200.h").read_text()
    declared = set(re.findall(r"\b(fsr_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fsr_abi_version() == 1
    import ctypes as C

    assert C.sizeof(_lib.TileParams) == 17 * 4  # layout of fsr_tile_params


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under floodsr_b200/ may import it."""
    for py in (REPO / "floodsr_b200").rglob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), py


def test_engine_fails_loudly_without_gpu(h1_model_fp):
    """No CPU fallback: creating the engine on a machine without a CUDA device is an error."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from floodsr_b200.engine import EngineB200

    with pytest.raises(_lib.EngineLibraryError, match="no CUDA device"):
        EngineB200(h1_model_fp)


def test_engine_base_is_abstract_and_dummy_subclass_works():
    """Mirror of the reference's tests/test_engine_contracts.py:30-60."""
    from floodsr_b200.engine import EngineBase, get_onnxruntime_info, get_rasterio_info

    with pytest.raises(TypeError):
        EngineBase()

    class DummyEngine(EngineBase):
        def __init__(self):
            self._model_fp = Path("dummy.onnx")

        def load(self) -> None:
            pass

        def run_tile(self, depth_lr_m, dem_hr_m, **kwargs):
            return {"prediction_m": np.asarray(dem_hr_m, dtype=np.float32)}

        def model_path(self) -> Path:
            return self._model_fp

    res = DummyEngine().run_tile(np.zeros((2, 2), np.float32), np.ones((2, 2), np.float32))
    assert res["prediction_m"].dtype == np.float32 and res["prediction_m"].size > 0
    for getter, key in [(get_onnxruntime_info, "available_providers"), (get_rasterio_info, "version")]:
        info = getter()
        assert isinstance(info.get("installed"), bool) and key in info
