"""Grid change around the tile loop (SURVEY.md section 8 f#2): oracle properties on CPU, CUDA kernel vs oracle on the GPU.

The reference reaches this arithmetic through rasterio.warp.reproject (GDAL), which is absent here: the oracle restates
GDAL's bilinear kernels (oracle/resample_np.py, "parity unpinned") and the CUDA path must agree with it bit for bit.
"""
import numpy as np
import pytest

from floodsr_b200.resample import bounds_to_transform
from oracle.resample_np import resample_bilinear as oracle_resample


def _grid(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    return (300.0 + 40.0 * np.sin(xx / 17.0) * np.cos(yy / 23.0) + rng.normal(0.0, 0.5, (h, w))).astype(np.float32)


# (source shape, destination shape) spanning the same bounds: the reference's 15x -> 16x DEM case and its way back,
# plain 2x up / down, anisotropic, identity
CASES = [((225, 240), (240, 256)), ((240, 256), (225, 240)), ((64, 48), (128, 96)), ((128, 96), (64, 48)),
         ((90, 200), (96, 150)), ((50, 70), (50, 70))]


def _transforms(src_shape, dst_shape, bounds=(1000.0, 5000.0, 1960.0, 5900.0)):
    west, south, east, north = bounds
    return (bounds_to_transform(west, south, east, north, src_shape[1], src_shape[0]),
            bounds_to_transform(west, south, east, north, dst_shape[1], dst_shape[0]))


def test_bounds_to_transform_matches_affine_from_bounds():
    # rasterio.transform.from_bounds == Affine.translation(west, north) * Affine.scale((east-west)/w, (south-north)/h)
    t = bounds_to_transform(10.0, 20.0, 110.0, 70.0, 50, 25)
    assert t == (2.0, 0.0, 10.0, 0.0, -2.0, 70.0)


def test_oracle_identity_and_constant():
    src = _grid(40, 60, 0)
    ts, _ = _transforms(src.shape, src.shape)
    assert np.array_equal(oracle_resample(src, ts, src.shape, ts), src)
    for s_shape, d_shape in CASES:
        ts, td = _transforms(s_shape, d_shape)
        out = oracle_resample(np.full(s_shape, 7.25, np.float32), ts, d_shape, td)
        assert np.array_equal(out, np.full(d_shape, 7.25, np.float32)), (s_shape, d_shape)


def test_oracle_reproduces_a_plane_when_upsampling():
    s_shape, d_shape = (225, 240), (240, 256)
    ts, td = _transforms(s_shape, d_shape)
    yy, xx = np.mgrid[0:s_shape[0], 0:s_shape[1]]
    src = (2.0 * xx + 3.0 * yy + 1.0).astype(np.float32)
    out = oracle_resample(src, ts, d_shape, td)
    sx = (np.arange(d_shape[1]) + 0.5) * s_shape[1] / d_shape[1] - 0.5
    sy = (np.arange(d_shape[0]) + 0.5) * s_shape[0] / d_shape[0] - 0.5
    want = 2.0 * sx[None, :] + 3.0 * sy[:, None] + 1.0
    assert np.abs(out[2:-2, 2:-2] - want[2:-2, 2:-2]).max() < 1e-3


@pytest.mark.parametrize("s_shape,d_shape", [((225, 240), (240, 256)), ((64, 48), (128, 96)), ((32, 32), (512, 512)), ((50, 70), (50, 70))])
def test_oracle_up_sampling_agrees_with_opencv_and_scipy(s_shape, d_shape):
    """Independent cross-check of the 4-sample path (the reference's 15x -> 16x DEM case and the 16x way back are up-sampling).

    GDAL is absent, but two unrelated libraries implement the same published convention -- sample at pixel centres, weights
    1 - |distance|, at the raster edge only the pixels that exist count (weight renormalisation == border replication for
    two taps): `cv2.resize(INTER_LINEAR)` and `scipy.ndimage.map_coordinates(order=1, mode="nearest")`.  This pins the
    restatement's geometry and edge rule; the down-sampling kernel (widened triangle filter) stays GDAL-specific and
    unpinned."""
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    src = _grid(*s_shape, seed=11)
    ts, td = _transforms(s_shape, d_shape)
    got = oracle_resample(src, ts, d_shape, td)
    want_cv = cv2.resize(src, (d_shape[1], d_shape[0]), interpolation=cv2.INTER_LINEAR)
    sy = (np.arange(d_shape[0]) + 0.5) * s_shape[0] / d_shape[0] - 0.5
    sx = (np.arange(d_shape[1]) + 0.5) * s_shape[1] / d_shape[1] - 0.5
    want_sp = ndi.map_coordinates(src.astype(np.float64), np.meshgrid(sy, sx, indexing="ij"), order=1, mode="nearest")
    scale = float(np.abs(src).max())
    assert np.abs(got - want_sp).max() <= 2e-7 * scale      # float64 accumulation on both sides, one float32 rounding
    assert np.abs(got - want_cv).max() <= 2e-5 * scale      # OpenCV interpolates with 11-bit fixed-point weights for some sizes


def test_oracle_nodata_and_outside():
    s_shape, d_shape = (60, 75), (64, 80)
    ts, td = _transforms(s_shape, d_shape)
    src = _grid(*s_shape, 1)
    src[10:20, 30:45] = -9999.0
    out = oracle_resample(src, ts, d_shape, td, -9999.0, -9999.0)
    hole = out == -9999.0
    assert 0 < hole.sum() < 11 * 17 * 1.3          # only destinations with no valid neighbour keep nodata
    assert out[~hole].min() > 200.0                # nodata never bleeds into valid values
    # destination grid larger than the source bounds: pixels whose centre lies outside keep the fill value
    td_big = bounds_to_transform(900.0, 4900.0, 2060.0, 6000.0, 100, 100)
    big = oracle_resample(src, ts, (100, 100), td_big, None, -1.0)
    assert big[0, 0] == -1.0 and big[-1, -1] == -1.0 and big[50, 50] != -1.0


@pytest.fixture(scope="module")
def gpu_engine(tmp_path_factory):
    from floodsr_b200.engine import EngineB200
    from floodsr_b200.h1 import write_h1_model

    fp = write_h1_model(tmp_path_factory.mktemp("resample") / "model_infer.onnx", seed=0)
    eng = EngineB200(fp)
    eng.load()
    yield eng
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("s_shape,d_shape", CASES)
def test_cuda_resample_is_bit_identical_to_oracle(gpu_engine, s_shape, d_shape):
    from floodsr_b200.resample import resample_bilinear

    ts, td = _transforms(s_shape, d_shape)
    src = _grid(*s_shape, s_shape[0])
    assert np.array_equal(resample_bilinear(gpu_engine, src, ts, d_shape, td), oracle_resample(src, ts, d_shape, td))
    src[5:17, 8:30] = -9999.0
    got = resample_bilinear(gpu_engine, src, ts, d_shape, td, -9999.0, -9999.0)
    assert np.array_equal(got, oracle_resample(src, ts, d_shape, td, -9999.0, -9999.0))
    # shifted, larger destination: fill outside the source
    td_big = bounds_to_transform(930.0, 4950.0, 2060.0, 5990.0, d_shape[1] + 9, d_shape[0] + 5)
    shape_big = (d_shape[0] + 5, d_shape[1] + 9)
    got = resample_bilinear(gpu_engine, src, ts, shape_big, td_big, -9999.0, -5.0)
    assert np.array_equal(got, oracle_resample(src, ts, shape_big, td_big, -9999.0, -5.0))


@pytest.mark.gpu
def test_cuda_resample_full_size_reference_case(gpu_engine):
    """BASELINE config 2 (mersch-like): raw DEM 3840 x 3840 (15x) -> model grid 4096 x 4096 and the prediction back."""
    from floodsr_b200.resample import align_dem_to_model_grid, prediction_to_raw_grid

    raw = _grid(3840, 3840, 7)
    bounds = (0.0, 0.0, 7680.0, 7680.0)
    t_raw = bounds_to_transform(*bounds, 3840, 3840)
    got = align_dem_to_model_grid(gpu_engine, raw, t_raw, bounds, (256, 256), 16, dem_nodata=-9999.0)
    assert got["dem_hr"].shape == (4096, 4096) and got["resampled"] and got["crop_shape"] == (4096, 4096)
    want = oracle_resample(raw, t_raw, (4096, 4096), got["dem_hr_transform"], -9999.0, -9999.0)
    assert np.array_equal(got["dem_hr"], want)
    back = prediction_to_raw_grid(gpu_engine, got["dem_hr"], got["dem_hr_transform"], (3840, 3840), t_raw)
    assert np.array_equal(back, oracle_resample(got["dem_hr"], got["dem_hr_transform"], (3840, 3840), t_raw))
    assert np.abs(back - raw)[8:-8, 8:-8].max() < 2.5   # smoothing of the noise term only: the terrain survives the round trip
    same = prediction_to_raw_grid(gpu_engine, raw, t_raw, (3840, 3840), t_raw)
    assert same is not None and np.array_equal(same, raw)
