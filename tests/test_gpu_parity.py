"""Parity of the CUDA path against the CPU oracle, through the C ABI (run on the B200 box: -m gpu).

Bit-exact: window geometry, DEM statistics and normalisation, mosaic arithmetic.  Tolerance-level: anything
through log1p/expm1 (libm vs CUDA differ in the last ulp) and the network forward (fp32 <= 1e-4 m).

`precision="fp32"` is the fp32-tolerance mode on tcgen05 (split fp16 operands, three MMAs per product); the plain
fp32 CUDA-core kernels stay available as `precision="fp32_simt"` and are used here as a second, independent check.
"""

from __future__ import annotations

import importlib.util
from pathlib import Path

import numpy as np
import pytest

from floodsr_b200.synth import synth_dem, synth_depth, synth_raster, synth_tile

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

FP32_TOL_M = 1e-4  # north star: fp32 predictions within 1e-4 m of the reference CPU path


@pytest.fixture(scope="module")
def engine(h1_model_fp):
    from floodsr_b200.engine import EngineB200

    eng = EngineB200(h1_model_fp, precision="fp32")
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def oracle_engine(h1_model_fp):
    from oracle.engine_ref import OracleEngine

    return OracleEngine(h1_model_fp)


def _hex(stats):
    return {k: float(v).hex() for k, v in stats.items()}


# ---------------------------------------------------------------------------------------------------------
# a5-a8: per-tile normalisation
# ---------------------------------------------------------------------------------------------------------


@pytest.fixture(params=["0", "1"], ids=["one-cta", "three-kernel"])
def norm_path(request, monkeypatch):
    """Both normalisation paths of launch_tile_normalize (k_prologue.cu): one CTA per tile (large batches) and the
    scan / select / apply kernels (batches of <= 2 tiles per SM).  FSR_NORM_SPLIT forces one or the other."""
    monkeypatch.setenv("FSR_NORM_SPLIT", request.param)
    return request.param


@pytest.mark.parametrize("case", mg.tile_cases(), ids=lambda c: c[0])
def test_normalize_stage_matches_reference_golden_bit_exact(engine, golden, case, norm_path):
    from oracle import preprocessing_np as pp

    meta, arrays = golden
    name, depth, dem, kw = case
    g = meta["cases"][f"pre/{name}"]
    got = engine.stage_normalize(depth[None], dem[None], **kw)
    stats = {"p_clip": float(got["stats"][0, 0]), "dem_min": float(got["stats"][0, 1]), "dem_max": float(got["stats"][0, 2])}
    assert _hex(stats) == g["stats"]  # exact np.nanpercentile / min / max
    assert mg.digest(got["dem_norm"][0]) == g["dem_norm_sha"]  # bit-exact normalised DEM
    want_depth = pp.scale_depth_log1p(pp.replace_nodata_with_zero(depth, kw.get("depth_lr_nodata")), kw.get("max_depth", 5.0))
    assert np.abs(got["depth_norm"][0] - want_depth).max() <= 2.4e-7  # log1pf vs libm: <= 2 ulp at 1.0


def test_normalize_stage_random_tiles_vs_oracle(engine, norm_path):
    from oracle import preprocessing_np as pp

    rng = np.random.default_rng(5)
    b = 6
    dem = np.stack([synth_dem(512, 512, seed=100 + i) * np.float32(rng.choice([1.0, 1e-3, 7.5])) for i in range(b)])
    dem[1, 100:, :] = 0.0  # zero padding
    dem[2] = np.round(dem[2], 0)  # heavy ties
    depth = np.stack([synth_depth(32, 32, seed=100 + i) for i in range(b)])
    for pct in (95.0, 37.5, 100.0):
        got = engine.stage_normalize(depth, dem, dem_pct_clip=pct)
        for i in range(b):
            want, st = pp.normalize_dem(dem[i], pct)
            assert (float(got["stats"][i, 0]), float(got["stats"][i, 1]), float(got["stats"][i, 2])) == (
                st["p_clip"], st["dem_min"], st["dem_max"]), (pct, i)
            assert np.array_equal(got["dem_norm"][i], want), (pct, i)


def test_normalize_stage_ref_stats_and_errors(engine, norm_path):
    from oracle import preprocessing_np as pp

    depth, dem = synth_tile(9)
    ref = {"p_clip": 900.0, "dem_min": 250.0, "dem_max": 900.0}
    got = engine.stage_normalize(depth[None], dem[None], dem_ref_stats=ref)
    assert np.array_equal(got["dem_norm"][0], pp.normalize_dem(dem, ref_stats=ref)[0])
    with pytest.raises(AssertionError, match=r"DEM range must be > 0; got min=7.0, max=7.0"):
        engine.stage_normalize(depth[None], np.full((1, 512, 512), 7.0, np.float32))
    bad = dem.copy()
    bad[17, 33] = np.nan
    with pytest.raises(AssertionError, match="DEM contains non-finite values after nodata replacement"):
        engine.stage_normalize(depth[None], bad[None])
    badd = depth.copy()
    badd[3, 3] = np.inf
    with pytest.raises(AssertionError, match="low-res depth contains non-finite values after nodata replacement"):
        engine.stage_normalize(badd[None], dem[None])
    with pytest.raises(AssertionError, match="dem_pct_clip must be finite"):
        engine.stage_normalize(depth[None], dem[None], dem_pct_clip=0.0)
    # the engine is still usable after a failed call
    assert np.array_equal(engine.stage_normalize(depth[None], dem[None])["dem_norm"][0], pp.normalize_dem(dem)[0])


# ---------------------------------------------------------------------------------------------------------
# a11, a15: invert + mosaic
# ---------------------------------------------------------------------------------------------------------


def test_invert_stage_close_to_numpy(engine):
    from oracle import preprocessing_np as pp

    x = np.linspace(-0.2, 1.2, 200_001, dtype=np.float32)
    for d in (5.0, 3.0):
        got = engine.stage_invert(x, d)
        want = pp.invert_depth_log1p(x, d)
        assert np.abs(got - want).max() <= 2e-6  # expm1f vs libm expm1 on values up to D
        assert got.min() == 0.0 and got.max() == np.float32(d)


class _ReplayEngine:
    """Feeds pre-computed tiles to the oracle's tile loop in window order."""

    def __init__(self, tiles):
        self.tiles, self.i = tiles, 0

    def run_tile(self, depth, dem, **kw):
        t = self.tiles[self.i]
        self.i += 1
        return {"prediction_m": t, "dem_stats_used": {"p_clip": 1.0, "dem_min": 0.0, "dem_max": 1.0}}


@pytest.mark.parametrize(
    "h,w,method,ov",
    [(1024, 1536, "feather", 8), (1024, 1536, "hard", 8), (976, 1104, "feather", 8), (976, 1104, "feather", 4),
     (528, 1296, "feather", 12), (512, 512, "feather", 8), (1000, 1100, "feather", 8), (2000, 600, "hard", 8)],
)
def test_blend_stage_bit_exact_vs_oracle(engine, h, w, method, ov):
    from floodsr_b200.tiling import window_grid
    from oracle.stitch_np import run_tiled

    ys, xs = window_grid(h, w, 512, method, ov * 16)
    rng = np.random.default_rng(h * 7 + w)
    tiles = (rng.random((len(ys) * len(xs), 512, 512), dtype=np.float32) * 6.0 - 0.5).astype(np.float32)
    depth = np.zeros((h // 16, w // 16), np.float32)
    dem = np.zeros((h, w), np.float32)
    want, n_tiles, _ = run_tiled(_ReplayEngine(tiles), depth, dem, window_method=method, overlap_lr=ov)
    got = engine.stage_blend(tiles, h, w, method, ov)
    assert n_tiles == len(tiles)
    assert got.shape == want.shape and np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------
# a10: network forward
# ---------------------------------------------------------------------------------------------------------


def test_forward_stage_fp32_close_to_oracle(engine, oracle_engine):
    from oracle import preprocessing_np as pp

    b = 3
    dn = np.stack([pp.scale_depth_log1p(synth_depth(32, 32, seed=20 + i), 5.0) for i in range(b)])
    en = np.stack([pp.normalize_dem(synth_dem(512, 512, seed=20 + i))[0] for i in range(b)])
    want = oracle_engine.forward_norm(dn, en)
    got = engine.stage_forward(dn, en)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() <= 1e-5  # normalised units; x10.75 m/unit worst-case gain -> 1e-4 m
    assert want.std() > 0.02


def test_fp32_mode_runs_on_tensor_cores_and_agrees_with_the_cuda_core_kernels(engine, h1_model_fp, oracle_engine, monkeypatch):
    """The <= 1e-4 m mode three ways: fused split kernel, split LR layers + fp32 FMA high-resolution pair, all fp32 FMA."""
    from floodsr_b200.engine import EngineB200
    from oracle import preprocessing_np as pp

    b = 5  # not a multiple of anything: ragged row ranges per SM
    dn = np.stack([pp.scale_depth_log1p(synth_depth(32, 32, seed=30 + i), 5.0) for i in range(b)])
    en = np.stack([pp.normalize_dem(synth_dem(512, 512, seed=30 + i))[0] for i in range(b)])
    want = oracle_engine.forward_norm(dn[:2], en[:2])
    got = engine.stage_forward(dn, en)
    assert np.abs(got[:2] - want).max() <= 1e-5
    simt = EngineB200(h1_model_fp, precision="fp32_simt")
    ref = simt.stage_forward(dn, en)
    simt.close()
    assert np.abs(ref[:2] - want).max() <= 1e-5
    assert np.abs(got - ref).max() <= 1e-5
    monkeypatch.setenv("FSR_X3_HR_SIMT", "1")
    pair = EngineB200(h1_model_fp, precision="fp32")
    monkeypatch.delenv("FSR_X3_HR_SIMT")
    mid = pair.stage_forward(dn, en)
    # the low-resolution result is identical (same kernels); only the high-resolution layers differ
    assert np.array_equal(pair.debug_tensor(37, b), engine.debug_tensor(37, b))
    pair.close()
    assert np.abs(mid - got).max() <= 5e-6
    assert np.array_equal(engine.stage_forward(dn, en), got)  # deterministic


# ---------------------------------------------------------------------------------------------------------
# a1-a4: the engine contract (mirror of the reference's tests/test_engine_contracts.py:63-93)
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("repeat_run", [False, True], ids=["b200_contract_single_run", "b200_contract_repeat_run_is_deterministic"])
def test_engine_b200_run_tile_contract(h1_model_fp, ort_tile_inputs, repeat_run):
    from floodsr_b200.engine import EngineB200

    eng = EngineB200(h1_model_fp)
    run1 = eng.run_tile(
        ort_tile_inputs["depth_lr"], ort_tile_inputs["dem_hr"],
        depth_lr_nodata=ort_tile_inputs["depth_lr_nodata"], dem_hr_nodata=ort_tile_inputs["dem_hr_nodata"],
    )
    assert run1["prediction_m"].dtype == np.float32
    assert run1["prediction_m"].size > 0
    assert set(run1) == {"prediction_m", "prediction_norm", "dem_stats_used", "runtime_s"}
    assert run1["dem_stats_used"] == {"p_clip": 975.0, "dem_min": 500.0, "dem_max": 975.0}
    if repeat_run:
        run2 = eng.run_tile(
            ort_tile_inputs["depth_lr"], ort_tile_inputs["dem_hr"],
            depth_lr_nodata=ort_tile_inputs["depth_lr_nodata"], dem_hr_nodata=ort_tile_inputs["dem_hr_nodata"],
        )
        assert isinstance(run2["prediction_m"], np.ndarray)
        assert np.array_equal(run1["prediction_m"], run2["prediction_m"])
    assert eng.model_path() == Path(h1_model_fp).resolve()
    assert (eng.contract.scale, eng.contract.depth_lr_hwc, eng.contract.dem_hr_hwc) == (16, (32, 32, 1), (512, 512, 1))
    eng.close()
    assert eng.session is None and eng.contract is None


@pytest.mark.parametrize("case", mg.tile_cases(), ids=lambda c: c[0])
def test_run_tile_matches_oracle(engine, oracle_engine, case):
    name, depth, dem, kw = case
    want = oracle_engine.run_tile(depth, dem, **kw)
    got = engine.run_tile(depth, dem, **kw)
    assert got["dem_stats_used"] == want["dem_stats_used"]
    assert np.abs(got["prediction_norm"] - want["prediction_norm"]).max() <= 1e-5
    assert np.abs(got["prediction_m"] - want["prediction_m"]).max() <= FP32_TOL_M
    assert np.array_equal(got["prediction_m"] > 0.01, want["prediction_m"] > 0.01) or (
        np.abs(got["prediction_m"] - want["prediction_m"])[(got["prediction_m"] > 0.01) != (want["prediction_m"] > 0.01)].max() < 1e-5
    )


def test_run_tile_prenormalised_branch_and_shape_errors(engine, oracle_engine):
    from oracle import preprocessing_np as pp

    depth, dem = synth_tile(0)
    dn, en = pp.scale_depth_log1p(depth, 5.0), pp.normalize_dem(dem)[0]
    want = oracle_engine.run_tile(dn, en, normalize_inputs=False)
    got = engine.run_tile(dn, en, normalize_inputs=False)
    assert got["dem_stats_used"] == want["dem_stats_used"] == {"p_clip": 95.0, "dem_min": 0.0, "dem_max": 1.0}
    assert np.abs(got["prediction_m"] - want["prediction_m"]).max() <= FP32_TOL_M
    with pytest.raises(AssertionError, match=r"DEM tile must be normalized to \[0, 1\]"):
        engine.run_tile(dn, dem, normalize_inputs=False)
    with pytest.raises(AssertionError, match=r"depth tensor shape \(16, 16, 1\) != expected \(32, 32, 1\)"):
        engine.run_tile(depth[:16, :16], dem)
    with pytest.raises(AssertionError, match=r"DEM tensor shape"):
        engine.run_tile(depth, dem[:256])


def test_run_tiles_batch_equals_single_tiles(engine):
    b = 5
    depth = np.stack([synth_depth(32, 32, seed=40 + i) for i in range(b)])
    dem = np.stack([synth_dem(512, 512, seed=40 + i) for i in range(b)])
    batch = engine.run_tiles(depth, dem)
    for i in range(b):
        one = engine.run_tile(depth[i], dem[i])
        assert np.array_equal(batch["prediction_m"][i], one["prediction_m"])
        assert batch["dem_stats_used"][i] == one["dem_stats_used"]


# ---------------------------------------------------------------------------------------------------------
# a15: whole raster
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("h,w,method,ov", [(1024, 1536, "feather", 8), (976, 1104, "feather", 8), (976, 1104, "hard", 8)])
def test_run_raster_matches_oracle_tile_loop(engine, oracle_engine, h, w, method, ov):
    from oracle.stitch_np import run_tiled

    depth, dem = synth_raster(h, w, seed=h + w)
    want, n_tiles, summary = run_tiled(oracle_engine, depth, dem, window_method=method, overlap_lr=ov)
    got, got_n, got_summary = engine.run_raster(depth, dem, window_method=method, overlap_lr=ov)
    assert got.shape == want.shape == (h, w) and got.dtype == np.float32
    assert got_n == n_tiles
    assert got_summary == summary  # per-tile DEM stats are bit-exact, so is their float32 summary
    assert np.abs(got - want).max() <= FP32_TOL_M
    assert float(got.min()) >= 0.0 and float(got.max()) <= 5.0


def test_run_raster_properties_at_mersch_size(engine):
    """BASELINE config 2 size (4096 x 4096 model space, 121 feather windows): size-independent properties."""
    h = w = 4096
    depth, dem = synth_raster(h, w, seed=77)
    out1, n_tiles, summary = engine.run_raster(depth, dem)
    assert n_tiles == 121 and summary["tile_count"] == 121.0
    assert out1.shape == (h, w) and np.isfinite(out1).all() and out1.min() >= 0.0 and out1.max() <= 5.0
    out2, _, _ = engine.run_raster(depth, dem)
    assert np.array_equal(out1, out2)  # deterministic: no atomics anywhere in the path
    # hard mosaic == the batched tiles pasted side by side
    hard, n_hard, _ = engine.run_raster(depth, dem, window_method="hard")
    assert n_hard == 64
    d_t = depth.reshape(8, 32, 8, 32).transpose(0, 2, 1, 3).reshape(64, 32, 32)
    e_t = dem.reshape(8, 512, 8, 512).transpose(0, 2, 1, 3).reshape(64, 512, 512)
    tiles = engine.run_tiles(d_t, e_t, want_norm=False)["prediction_m"]
    pasted = tiles.reshape(8, 8, 512, 512).transpose(0, 2, 1, 3).reshape(h, w)
    assert np.array_equal(hard, pasted)
    # interior of non-overlapped regions of the feather mosaic equals the hard tile there
    assert np.array_equal(out1[:384, :384], hard[:384, :384])


@pytest.mark.parametrize("precision", ["fp32"])
def test_in_kernel_dem_normalisation_is_bit_identical_to_the_materialised_tiles(h1_model_fp, monkeypatch, precision):
    """In the fp32 mode the fused high-resolution kernel normalises each tile's raster window itself (no dem_norm tensor in HBM;
    the 16-bit kernel's rows are too short to hide the divisions, that mode keeps the materialised tiles);
    FSR_NO_LAZY_DEM=1 keeps the normalisation kernel's materialised tiles.  Same operation sequence -> same bits: ragged
    raster (zero padding beyond the raster), nodata replacement, caller-supplied statistics with a negative minimum."""
    from floodsr_b200.engine import EngineB200

    h, w = 976, 1104
    depth, dem = synth_raster(h, w, seed=5)
    lazy = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.setenv("FSR_NO_LAZY_DEM", "1")
    mat = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.delenv("FSR_NO_LAZY_DEM")
    for method in ("feather", "hard"):
        a, na, sa = lazy.run_raster(depth, dem, window_method=method)
        b, nb, sb = mat.run_raster(depth, dem, window_method=method)
        assert na == nb and sa == sb and np.array_equal(a, b), method
    d_t = np.stack([synth_depth(32, 32, seed=70 + i) for i in range(3)])
    e_t = np.stack([synth_dem(512, 512, seed=70 + i) for i in range(3)])
    e_t[1, 100:140, 200:260] = -9999.0
    ref = {"p_clip": 900.0, "dem_min": -50.0, "dem_max": 900.0}
    for kw in ({}, {"dem_ref_stats": ref}, {"dem_pct_clip": 37.5}, {"dem_hr_nodata": -9999.0}):
        a = lazy.run_tiles(d_t, e_t, **kw)
        b = mat.run_tiles(d_t, e_t, **kw)
        assert np.array_equal(a["prediction_m"], b["prediction_m"]) and np.array_equal(a["prediction_norm"], b["prediction_norm"]), kw
        assert a["dem_stats_used"] == b["dem_stats_used"]
    lazy.close()
    mat.close()


# ---------------------------------------------------------------------------------------------------------
# graphs other than H1: operator spellings of a tf2onnx export, other widths (the loader runs whatever the file holds)
# ---------------------------------------------------------------------------------------------------------


def _engine_vs_oracle_on_model(fp, precision, tol_norm, seeds=(3, 4)):
    from floodsr_b200.engine import EngineB200
    from oracle import preprocessing_np as pp
    from oracle.engine_ref import OracleEngine

    dn = np.stack([pp.scale_depth_log1p(synth_depth(32, 32, seed=s), 5.0) for s in seeds])
    en = np.stack([pp.normalize_dem(synth_dem(512, 512, seed=s))[0] for s in seeds])
    want = OracleEngine(fp).forward_norm(dn, en)
    eng = EngineB200(fp, precision=precision)
    got = eng.stage_forward(dn, en)
    eng.close()
    assert want.std() > 1e-3
    err = float(np.abs(got - want).max())
    assert err <= tol_norm, (precision, err)
    return err


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fp16", 4e-3), ("fp32_simt", 1e-5)])
def test_export_spelling_variants_on_the_engine(tmp_path, precision, tol):
    """Pad + VALID conv, stride-2 convs (both padding conventions), bilinear Resize (half_pixel and align_corners), unfolded
    normalisation (Mul / Sub / Div), Clip, Sigmoid, Reshape / Cast -- all in one graph, against the oracle's interpreter."""
    from floodsr_b200.onnx_io import save_onnx
    from tests import h1_variants as V

    fp = tmp_path / "everything.onnx"
    save_onnx(V.everything(), fp)
    _engine_vs_oracle_on_model(fp, precision, tol)


@pytest.mark.parametrize("ctm", ["asymmetric"])
def test_bilinear_asymmetric_resize_on_the_engine(tmp_path, ctm):
    from floodsr_b200.h1 import build_h1_model
    from floodsr_b200.onnx_io import save_onnx
    from tests import h1_variants as V

    fp = tmp_path / "bilinear.onnx"
    save_onnx(V.bilinear_resize(build_h1_model(seed=6), ctm, which=(0, 1, 2, 3)), fp)
    _engine_vs_oracle_on_model(fp, "fp32", 1e-5)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_activation_overflow_is_reported_not_clipped_to_zero(tmp_path, precision):
    """Both tensor-core modes keep activations in fp16 (pairs): a network whose activations exceed 65504 produces Inf / NaN,
    which the final clip would turn into 0 m.  The engine raises instead (FSR_FLAG_PRED_NONFINITE)."""
    from floodsr_b200.engine import EngineB200
    from floodsr_b200.h1 import build_h1_model
    from floodsr_b200.onnx_io import save_onnx

    m = build_h1_model(seed=0)
    name = next(n.inputs[1] for n in m.nodes if n.op_type == "Conv" and m.initializers[n.inputs[1]].shape[:2] == (32, 32))
    m.initializers[name] = (m.initializers[name] * np.float32(3e5)).astype(np.float32)
    fp = tmp_path / "overflow.onnx"
    save_onnx(m, fp)
    eng = EngineB200(fp, precision=precision)
    depth, dem = synth_tile(1)
    with pytest.raises(FloatingPointError, match="non-finite"):
        eng.run_tile(depth, dem)
    eng.close()


@pytest.mark.parametrize("f", [16, 64])
def test_other_widths_run_with_the_fp32_high_resolution_pair(tmp_path, f):
    """base_filters 16 / 64: the low-resolution layers stay on tcgen05 (16-channel outputs, 1024-channel deep level), the
    high-resolution end (not 32 channels) runs on the fp32 FMA kernels -- in both tensor-core modes."""
    from floodsr_b200.h1 import build_h1_model
    from floodsr_b200.onnx_io import save_onnx

    fp = tmp_path / f"h1_f{f}.onnx"
    save_onnx(build_h1_model(seed=2, base_filters=f), fp)
    _engine_vs_oracle_on_model(fp, "fp32", 1e-5, seeds=(3,))
    _engine_vs_oracle_on_model(fp, "fp16", 4e-3, seeds=(3,))


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs 2-4 against the oracle's tile loop (oracle/stitch_np.run_tiled over the torch-CPU network)
# ---------------------------------------------------------------------------------------------------------


def _wet_mask_agrees(got, want, tol):
    flips = (got > 0.01) != (want > 0.01)
    err = np.abs(got - want)
    return (not flips.any()) or float(np.abs(want[flips] - 0.01).max()) <= float(err[flips].max()) <= tol


@pytest.mark.parametrize("method", ["feather", "hard"])
def test_config2_mersch_size_raster_vs_oracle(engine, oracle_engine, method):
    """BASELINE config 2: 4096 x 4096 model-space raster, 121 feather / 64 hard windows; fp32 mode <= 1e-4 m."""
    from oracle.stitch_np import run_tiled

    h = w = 4096
    depth, dem = synth_raster(h, w, seed=77)
    want, n_tiles, summary = run_tiled(oracle_engine, depth, dem, window_method=method, overlap_lr=8)
    got, got_n, got_summary = engine.run_raster(depth, dem, window_method=method)
    assert got_n == n_tiles == (121 if method == "feather" else 64) and got_summary == summary
    assert np.abs(got - want).max() <= FP32_TOL_M
    assert _wet_mask_agrees(got, want, FP32_TOL_M)


def test_config3_batch_of_256_distinct_tiles_vs_oracle(engine, oracle_engine, h1_model_fp):
    """BASELINE config 3: 256 independent tiles; fp32 mode <= 1e-4 m and fp16 mode <= 1e-2 m + wet/dry mask, every tile
    against the oracle's run_tile."""
    from floodsr_b200.engine import EngineB200

    n = 256
    depth = np.stack([synth_depth(32, 32, seed=500 + i) for i in range(n)])
    dem = np.stack([synth_dem(512, 512, seed=500 + i) for i in range(n)])
    want = [oracle_engine.run_tile(depth[i], dem[i]) for i in range(n)]
    got32 = engine.run_tiles(depth, dem, want_norm=False)
    fp16 = EngineB200(h1_model_fp, precision="fp16")
    got16 = fp16.run_tiles(depth, dem, want_norm=False)
    fp16.close()
    for i in range(n):
        assert got32["dem_stats_used"][i] == want[i]["dem_stats_used"] == got16["dem_stats_used"][i], i
        assert np.abs(got32["prediction_m"][i] - want[i]["prediction_m"]).max() <= FP32_TOL_M, i
        assert np.abs(got16["prediction_m"][i] - want[i]["prediction_m"]).max() <= 1e-2, i
        assert _wet_mask_agrees(got16["prediction_m"][i], want[i]["prediction_m"], 1e-2), i


def test_config4_8k_raster_vs_oracle(engine, oracle_engine, h1_model_fp):
    """BASELINE config 4: 8192 x 8192 raster (512 x 512 lores), 441 feather windows; fp32 mode <= 1e-4 m, fp16 <= 1e-2 m."""
    from floodsr_b200.engine import EngineB200
    from oracle.stitch_np import run_tiled

    h = w = 8192
    depth, dem = synth_raster(h, w, seed=88)
    want, n_tiles, summary = run_tiled(oracle_engine, depth, dem, window_method="feather", overlap_lr=8)
    got, got_n, got_summary = engine.run_raster(depth, dem)
    assert got_n == n_tiles == 441 and got_summary == summary
    assert np.abs(got - want).max() <= FP32_TOL_M
    fp16 = EngineB200(h1_model_fp, precision="fp16")
    got16, _, _ = fp16.run_raster(depth, dem)
    fp16.close()
    assert np.abs(got16 - want).max() <= 1e-2
    assert _wet_mask_agrees(got16, want, 1e-2)


@pytest.mark.parametrize("h,w,ov", [(1536, 1024, 12), (2048, 528, 20), (1600, 640, 17)])
def test_large_overlaps_three_window_rows_per_coordinate(h1_model_fp, oracle_engine, monkeypatch, h, w, ov):
    """overlap_lr >= 17 (or 12 with a forced trailing window at H = 1536: window rows [0, 320, 640, 960, 1024]) covers
    coordinates with three or more window rows.  With one window row per pipeline band (FSR_BAND_TILES=1) a band would own
    fewer rows than its incoming halo covers unless the planner merges window rows: the mosaic must still equal the oracle's."""
    from floodsr_b200.engine import EngineB200
    from oracle.stitch_np import run_tiled

    monkeypatch.setenv("FSR_BAND_TILES", "1")
    eng = EngineB200(h1_model_fp, precision="fp16")
    monkeypatch.delenv("FSR_BAND_TILES")
    depth, dem = synth_raster(h, w, seed=h + w)
    tiles_engine = EngineB200(h1_model_fp, precision="fp16")  # default band size: the reference for bit identity
    want_bits, n_ref, _ = tiles_engine.run_raster(depth, dem, overlap_lr=ov)
    got, n, _ = eng.run_raster(depth, dem, overlap_lr=ov)
    assert n == n_ref and np.array_equal(got, want_bits)
    want, n_tiles, _ = run_tiled(oracle_engine, depth, dem, window_method="feather", overlap_lr=ov)
    assert n == n_tiles and np.abs(got - want).max() <= 1e-2
    eng.close()
    tiles_engine.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_two_phase_pipeline_is_bit_identical_to_band_pipeline(h1_model_fp, monkeypatch, precision):
    """fsr_run_raster runs the low-resolution layers per GROUP of window rows and the fused high-resolution kernel per window
    row (two-phase pipeline, the fp32 mode's default; FSR_PHASES=1 / 0 forces it on / off); otherwise every band runs in one piece.  Per-window results do not depend on how the
    windows are batched, so the two mosaics, and the mosaics of other group sizes, are the same bits (7 window rows: groups
    of 1, 2, 4 rows by default)."""
    from floodsr_b200.engine import EngineB200

    depth, dem = synth_raster(2816, 1280, seed=77)
    eng = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.setenv("FSR_PHASES", "1")
    got, n, summary = eng.run_raster(depth, dem)
    monkeypatch.setenv("FSR_PHASES", "0")
    want, n_ref, summary_ref = eng.run_raster(depth, dem)
    eng.close()
    assert n == n_ref and n >= 21 and summary == summary_ref
    assert np.array_equal(got, want)
    monkeypatch.setenv("FSR_GROUP_TILES", "8")  # two window rows per group at most
    small = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.delenv("FSR_GROUP_TILES")
    monkeypatch.setenv("FSR_PHASES", "1")
    got_small, _, _ = small.run_raster(depth, dem, window_method="hard", overlap_lr=0)
    monkeypatch.setenv("FSR_PHASES", "0")
    want_small, _, _ = small.run_raster(depth, dem, window_method="hard", overlap_lr=0)
    small.close()
    assert np.array_equal(got_small, want_small)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_last_band_in_column_parts_is_bit_identical(h1_model_fp, monkeypatch, precision):
    """The last band of the two-phase pipeline runs in column parts (kernel of part k+1 next to the D2H copy of part k): 34
    window columns give 4 parts.  Same bits as the band pipeline, feather and hard windows."""
    from floodsr_b200.engine import EngineB200

    depth, dem = synth_raster(1280, 13184, seed=5)
    eng = EngineB200(h1_model_fp, precision=precision)
    for kw in ({}, {"window_method": "hard", "overlap_lr": 0}):
        monkeypatch.setenv("FSR_PHASES", "1")
        got, n, _ = eng.run_raster(depth, dem, **kw)
        monkeypatch.setenv("FSR_TAIL_PARTS", "1")
        one, _, _ = eng.run_raster(depth, dem, **kw)
        monkeypatch.delenv("FSR_TAIL_PARTS")
        monkeypatch.setenv("FSR_PHASES", "0")
        want, n_ref, _ = eng.run_raster(depth, dem, **kw)
        assert n == n_ref and n >= 3 * 26
        assert np.array_equal(got, want) and np.array_equal(one, want)
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_normalisation_paths_feed_the_network_the_same_bits(h1_model_fp, monkeypatch, precision):
    """The pooled low-resolution DEM (and in the 16-bit mode the normalised tiles) come out of either normalisation path with
    the same summation order: predictions are bit-identical."""
    from floodsr_b200.engine import EngineB200

    depth = np.stack([synth_depth(32, 32, seed=40 + i) for i in range(5)])
    dem = np.stack([synth_dem(512, 512, seed=40 + i) for i in range(5)])
    dem[3, :, 200:] = 0.0
    eng = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.setenv("FSR_NORM_SPLIT", "0")
    a = eng.run_tiles(depth, dem)
    monkeypatch.setenv("FSR_NORM_SPLIT", "1")
    b = eng.run_tiles(depth, dem)
    eng.close()
    assert a["dem_stats_used"] == b["dem_stats_used"]
    assert np.array_equal(a["prediction_m"], b["prediction_m"])


@pytest.mark.parametrize("precision,tol_m", [("fp32", 5e-5), ("fp16", 5e-3)])
def test_folded_upsampling_agrees_with_the_two_op_path(h1_model_fp, oracle_engine, monkeypatch, precision, tol_m):
    """The 2x nearest upsampling in front of the image-major decoder convolutions is folded into them (weights of the taps that
    share a source pixel summed per output parity); FSR_NO_FOLD_UP=1 runs upsampling and convolution as two ops.  Both stay
    within the mode's tolerance of the oracle and close to one another (sum of products vs product with the summed weight)."""
    from floodsr_b200.engine import EngineB200

    depth = np.stack([synth_depth(32, 32, seed=60 + i) for i in range(3)])
    dem = np.stack([synth_dem(512, 512, seed=60 + i) for i in range(3)])
    folded = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.setenv("FSR_NO_FOLD_UP", "1")
    two_op = EngineB200(h1_model_fp, precision=precision)
    monkeypatch.delenv("FSR_NO_FOLD_UP")
    a = folded.run_tiles(depth, dem)["prediction_m"]
    b = two_op.run_tiles(depth, dem)["prediction_m"]
    assert folded.launch_count() < two_op.launch_count()  # two upsampling launches fewer per pass
    folded.close()
    two_op.close()
    want = np.stack([oracle_engine.run_tile(depth[i], dem[i])["prediction_m"] for i in range(3)])
    bound = FP32_TOL_M if precision == "fp32" else 1e-2
    assert np.abs(a - want).max() <= bound and np.abs(b - want).max() <= bound
    assert np.abs(a - b).max() <= tol_m


def test_run_raster_input_assertions(engine):
    depth, dem = synth_raster(1024, 1024, seed=3)
    with pytest.raises(AssertionError, match="depth shape"):
        engine.run_raster(depth[:-1], dem)
    with pytest.raises(AssertionError, match="aligned DEM must be 2D"):
        engine.run_raster(depth, dem[None])
    bad = dem.copy()
    bad[5, 5] = np.inf
    with pytest.raises(AssertionError, match="aligned DEM contains non-finite values"):
        engine.run_raster(depth, bad)
    with pytest.raises(AssertionError, match="unsupported window_method"):
        engine.run_raster(depth, dem, window_method="soft")
    with pytest.raises(AssertionError, match="feather windowing requires overlap_lr > 0"):
        engine.run_raster(depth, dem, overlap_lr=0)


# ---------------------------------------------------------------------------------------------------------
# row-band sharding on one device: two bands run back to back must reproduce the single pass bit for bit
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("h,w,world", [(2048, 1024, 2), (1552, 1040, 3), (4096, 1024, 4)])
def test_band_sharding_is_bit_identical_to_single_pass(engine, h, w, world, split):
    import torch

    from floodsr_b200.dist import CudaBandExecutor, plan_bands

    depth, dem = synth_raster(h, w, seed=h)
    want, _, _ = engine.run_raster(depth, dem)
    plans, ys, xs = plan_bands(h, w, 512, "feather", 128, world)
    ex = CudaBandExecutor(engine, h, w, "feather", 128)
    dev = torch.device("cuda", engine.device)
    got = np.empty_like(want)
    halo = None
    for plan in plans:
        if plan.empty:
            continue
        r0 = plan.in_row0
        dem_band = torch.from_numpy(dem[r0 : r0 + plan.in_rows]).to(dev).contiguous()
        depth_band = torch.from_numpy(depth[r0 // 16 : (r0 + plan.in_rows + 15) // 16]).to(dev).contiguous()
        halo_out = ex.band_run(plan, depth_band, dem_band, r0)
        assert (halo is None) == (plan.halo_in_rows == 0)
        if split:
            # what run_band_step does under a process group: the rows below the shared ones first, the shared ones last
            k = plan.halo_in_rows
            rows = ex.band_finalize_rows(plan, None, None, k, plan.n_rows)
            if k > 0:
                ex.band_finalize_rows(plan, halo, rows, 0, k)
        else:
            rows = ex.band_finalize(plan, halo)
        ex.check_flags()
        got[plan.row0 : plan.row0 + plan.n_rows] = rows.cpu().numpy()
        halo = halo_out
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------
# the 16-bit tensor-core mode (tcgen05, fp16 operands, fp32 accumulate): north-star tolerance 1e-2 m and the same
# wet/dry mask at a 0.01 m threshold
# ---------------------------------------------------------------------------------------------------------

# fp16 operands meet the north-star bound of the 16-bit mode (1e-2 m).  bf16 operands (3 fewer mantissa bits) measured
# 2.4e-2 m on the H1 graph in round 1, i.e. outside the bound: that mode was retired instead of tested against a looser one.
TC_TOL_M = {"fp16": 1e-2}
TC_TOL_NORM = {"fp16": 2e-3}
WET_M = 0.01


@pytest.fixture(scope="module", params=["fp16"])
def tc_engine(request, h1_model_fp):
    from floodsr_b200.engine import EngineB200

    eng = EngineB200(h1_model_fp, precision=request.param)
    yield eng
    eng.close()


def _assert_tc_close(eng, got_m, want_m):
    tol = TC_TOL_M[eng.precision]
    err = np.abs(got_m - want_m)
    assert float(err.max()) <= tol, float(err.max())
    flips = (got_m > WET_M) != (want_m > WET_M)
    # a pixel may only change class when the oracle itself sits on the threshold (closer than the actual error there)
    assert not flips.any() or float(np.abs(want_m[flips] - WET_M).max()) <= float(err[flips].max()) <= tol
    assert flips.mean() <= 2e-3, flips.mean()


@pytest.mark.parametrize("case", mg.tile_cases(), ids=lambda c: c[0])
def test_tc_run_tile_matches_oracle(tc_engine, oracle_engine, case):
    name, depth, dem, kw = case
    want = oracle_engine.run_tile(depth, dem, **kw)
    got = tc_engine.run_tile(depth, dem, **kw)
    assert got["dem_stats_used"] == want["dem_stats_used"]  # normalisation statistics stay bit-exact in every mode
    _assert_tc_close(tc_engine, got["prediction_m"], want["prediction_m"])
    assert np.abs(got["prediction_norm"] - want["prediction_norm"]).max() <= TC_TOL_NORM[tc_engine.precision]


@pytest.mark.parametrize("b", [1, 3, 7])
def test_tc_batches_split_rows_evenly_and_match_fp32_engine(tc_engine, engine, b):
    """Any batch size: the head kernel splits the batch's strip rows into equal ranges per SM (ragged items)."""
    depth = np.stack([synth_depth(32, 32, seed=60 + i) for i in range(b)])
    dem = np.stack([synth_dem(512, 512, seed=60 + i) for i in range(b)])
    want = engine.run_tiles(depth, dem)
    got = tc_engine.run_tiles(depth, dem)
    again = tc_engine.run_tiles(depth, dem)
    assert np.array_equal(got["prediction_m"], again["prediction_m"])  # deterministic
    assert got["dem_stats_used"] == want["dem_stats_used"]
    for i in range(b):
        _assert_tc_close(tc_engine, got["prediction_m"][i], want["prediction_m"][i])
        one = tc_engine.run_tile(depth[i], dem[i])
        assert np.array_equal(one["prediction_m"], got["prediction_m"][i])  # independent of the batch it ran in


def test_tc_run_raster_matches_oracle_tile_loop(tc_engine, oracle_engine):
    from oracle.stitch_np import run_tiled

    h, w = 976, 1104
    depth, dem = synth_raster(h, w, seed=h + w)
    want, n_tiles, summary = run_tiled(oracle_engine, depth, dem, window_method="feather", overlap_lr=8)
    got, got_n, got_summary = tc_engine.run_raster(depth, dem)
    assert got_n == n_tiles and got_summary == summary
    _assert_tc_close(tc_engine, got, want)


def test_fused_and_unfused_high_resolution_paths_agree(h1_model_fp, monkeypatch):
    """The one-kernel convT+head path (k_tc_fused.cu) against the convT -> HBM -> head kernel pair on the same tiles."""
    from floodsr_b200.engine import EngineB200

    b = 6
    depth = np.stack([synth_depth(32, 32, seed=80 + i) for i in range(b)])
    dem = np.stack([synth_dem(512, 512, seed=80 + i) for i in range(b)])
    fused = EngineB200(h1_model_fp, precision="fp16")
    monkeypatch.setenv("FSR_NO_FUSED_HR", "1")
    pair = EngineB200(h1_model_fp, precision="fp16")
    monkeypatch.delenv("FSR_NO_FUSED_HR")
    a = fused.run_tiles(depth, dem)
    c = pair.run_tiles(depth, dem)
    assert a["dem_stats_used"] == c["dem_stats_used"]
    # identical 16-bit feature values and weights; only the fp32 accumulation order inside the tensor core differs
    assert np.abs(a["prediction_norm"] - c["prediction_norm"]).max() <= 2e-6
    assert np.abs(a["prediction_m"] - c["prediction_m"]).max() <= 2e-5
    fused.close()
    pair.close()


@pytest.mark.parametrize("h,w,world", [(2048, 1024, 2), (4096, 1024, 3), (1664, 6400, 2)])  # 17 window columns: the last band runs in column parts
@pytest.mark.parametrize("phases", ["0", "1"])
def test_band_host_pipeline_is_bit_identical_to_single_pass(tc_engine, h, w, world, phases, monkeypatch):
    """Host-buffer band path (pipelined copies, deferred blend of the rows shared with the previous rank), as the band
    pipeline and as the two-phase pipeline (the fp32 mode's default)."""
    from floodsr_b200 import _lib
    from floodsr_b200.dist import CudaBandExecutor, plan_bands

    depth, dem = synth_raster(h, w, seed=h + 1)
    monkeypatch.setenv("FSR_PHASES", "0")
    want, _, _ = tc_engine.run_raster(depth, dem)
    monkeypatch.setenv("FSR_PHASES", phases)
    plans, ys, xs = plan_bands(h, w, 512, "feather", 128, world)
    ex = CudaBandExecutor(tc_engine, h, w, "feather", 128)
    got = np.empty_like(want)
    halo = None
    for plan in plans:
        if plan.empty:
            continue
        r0 = plan.in_row0
        dem_band = _lib.pinned_empty((plan.in_rows, w))
        dem_band[:] = dem[r0 : r0 + plan.in_rows]
        lr0, lr1 = r0 // 16, min((r0 + plan.in_rows + 15) // 16, h // 16)
        depth_band = _lib.pinned_empty((lr1 - lr0, w // 16))
        depth_band[:] = depth[lr0:lr1]
        out_rows = _lib.pinned_empty((plan.n_rows, w))
        halo_out = ex.band_host_begin(plan, depth_band, dem_band, r0, out_rows)
        ex.band_host_end(halo)
        got[plan.row0 : plan.row0 + plan.n_rows] = out_rows
        halo = halo_out
    assert np.array_equal(got, want)


@pytest.mark.parametrize("h,w,method", [(160, 208, "feather"), (528, 1296, "hard"), (1024, 512, "feather")])
def test_tc_run_raster_ragged_and_hard(tc_engine, oracle_engine, h, w, method):
    """Rasters smaller than / not a multiple of the tile (virtual zero padding) and the hard mosaic in the 16-bit modes."""
    from oracle.stitch_np import run_tiled

    depth, dem = synth_raster(h, w, seed=h * 3 + w)
    want, n_tiles, summary = run_tiled(oracle_engine, depth, dem, window_method=method, overlap_lr=8)
    got, got_n, got_summary = tc_engine.run_raster(depth, dem, window_method=method)
    assert got.shape == (h, w) and got_n == n_tiles and got_summary == summary
    _assert_tc_close(tc_engine, got, want)


# ---------------------------------------------------------------------------------------------------------
# BASELINE sizes in the tensor-core mode the bench runs (fp16 operands): size-independent properties
# ---------------------------------------------------------------------------------------------------------


@pytest.fixture(scope="module")
def fp16_engine(h1_model_fp):
    from floodsr_b200.engine import EngineB200

    eng = EngineB200(h1_model_fp, precision="fp16")
    yield eng
    eng.close()


def test_bench_size_raster_is_periodic_for_periodic_inputs(fp16_engine):
    """bench.py's workload size (4096 x 32768, 935 feather windows).  The inputs repeat every 384 columns (the window
    stride), so every interior window column sees the same pixels: its predictions, and therefore the mosaic, must repeat
    bit for bit with the same period — whatever chunk, image block or CTA a window lands in."""
    h, w, period = 4096, 32768, 384
    depth_p, dem_p = synth_raster(h, period, seed=21)
    reps = -(-w // period)
    dem = np.ascontiguousarray(np.tile(dem_p, (1, reps))[:, :w])
    depth = np.ascontiguousarray(np.tile(depth_p, (1, reps))[:, : w // 16])
    out, n_tiles, _ = fp16_engine.run_raster(depth, dem)
    assert n_tiles == 935 and out.shape == (h, w) and np.isfinite(out).all() and out.min() >= 0.0 and out.max() <= 5.0
    ref = out[:, 2 * period: 3 * period]
    for k in (3, 10, 41, 60, 82):   # windows k-1 and k are interior (not the first / forced last window column)
        assert np.array_equal(out[:, k * period: (k + 1) * period], ref), k
    again, _, _ = fp16_engine.run_raster(depth, dem)
    assert np.array_equal(out, again)


def test_bench_size_raster_fp32_mode_two_phase_pipeline(engine, monkeypatch):
    """The same workload in the fp32 mode, whose default is the two-phase pipeline (groups of 1, 2, 4, 4 window rows, last band
    in 4 column parts at 85 window columns): periodic like the inputs, and the same bits as the band pipeline."""
    h, w, period = 4096, 32768, 384
    depth_p, dem_p = synth_raster(h, period, seed=22)
    reps = -(-w // period)
    dem = np.ascontiguousarray(np.tile(dem_p, (1, reps))[:, :w])
    depth = np.ascontiguousarray(np.tile(depth_p, (1, reps))[:, : w // 16])
    monkeypatch.delenv("FSR_PHASES", raising=False)
    out, n_tiles, _ = engine.run_raster(depth, dem)
    assert n_tiles == 935 and out.shape == (h, w) and np.isfinite(out).all() and out.min() >= 0.0 and out.max() <= 5.0
    ref = out[:, 2 * period: 3 * period]
    for k in (3, 21, 22, 42, 43, 63, 64, 82):   # includes the window columns either side of the column-part cuts
        assert np.array_equal(out[:, k * period: (k + 1) * period], ref), k
    monkeypatch.setenv("FSR_PHASES", "0")
    bands, _, _ = engine.run_raster(depth, dem)
    assert np.array_equal(out, bands)


def test_batch_256_equals_single_tiles_in_tensor_core_mode(fp16_engine):
    """BASELINE config 3 (256 independent tiles): a tile's result does not depend on the batch it travels in."""
    from floodsr_b200.synth import synth_tile

    base = [synth_tile(s) for s in range(8)]
    depth = np.stack([base[i % 8][0] for i in range(256)])
    dem = np.stack([base[i % 8][1] for i in range(256)])
    for i in range(8, 256):                      # make the tiles distinct: shift the terrain, scale the depth
        dem[i] = dem[i] + np.float32(0.37 * i)
        depth[i] = depth[i] * np.float32(1.0 + 0.001 * i)
    batch = fp16_engine.run_tiles(depth, dem, want_norm=False)
    for i in (0, 7, 8, 100, 127, 128, 255):
        one = fp16_engine.run_tile(depth[i], dem[i])
        assert np.array_equal(batch["prediction_m"][i], one["prediction_m"]), i
        assert batch["dem_stats_used"][i] == one["dem_stats_used"]


@pytest.mark.skipif(not __import__("os").environ.get("FSR_BIG_TESTS"), reason="needs ~14 GB of host memory; set FSR_BIG_TESTS=1")
def test_32k_raster_on_one_gpu_is_periodic_in_both_axes(fp16_engine):
    """BASELINE config 5 (32768 x 32768, 7225 feather windows) through ONE engine: byte offsets beyond 2^32.  Inputs
    repeat every 384 pixels along both axes, so interior window rows and columns must repeat bit for bit."""
    n, period = 32768, 384
    depth_p, dem_p = synth_raster(period, period, seed=33)
    reps = -(-n // period)
    dem = np.ascontiguousarray(np.tile(dem_p, (reps, reps))[:n, :n])
    depth = np.ascontiguousarray(np.tile(depth_p, (reps, reps))[: n // 16, : n // 16])
    out, n_tiles, _ = fp16_engine.run_raster(depth, dem)
    assert n_tiles == 7225 and out.shape == (n, n)
    ref = out[2 * period: 3 * period, 2 * period: 3 * period]
    assert np.isfinite(ref).all() and ref.max() > 0.0
    for ky in (3, 30, 57, 82):           # rows beyond 2^32 bytes into the raster from ky = 43 on
        for kx in (2, 41, 83):
            assert np.array_equal(out[ky * period: (ky + 1) * period, kx * period: (kx + 1) * period], ref), (ky, kx)
