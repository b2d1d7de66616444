"""GeoTIFF in / out without rasterio and the file-to-file run (SURVEY.md section 8 f#3)."""
import numpy as np
import pytest

from floodsr_b200.raster_io import (TAG_GEOKEYS, bounds_of, clip_to_bounds, read_geotiff, window_from_bounds, write_geotiff)
from floodsr_b200.resample import bounds_to_transform
from floodsr_b200.synth import synth_raster

UTM32 = (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32632)       # projected, PixelIsArea, EPSG:32632
UTM32_POINT = (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 2, 3072, 0, 1, 32632)  # the same, PixelIsPoint


@pytest.mark.parametrize("compression", ["tiff_lzw", "raw", "tiff_adobe_deflate"])
def test_geotiff_round_trip(tmp_path, compression):
    arr = (np.random.default_rng(0).random((300, 500)) * 100).astype(np.float32)
    t = bounds_to_transform(1000.0, 5000.0, 2000.0, 5600.0, 500, 300)
    fp = write_geotiff(tmp_path / "a.tif", arr, t, nodata=-9999.0, geo_tags={TAG_GEOKEYS: UTM32}, compression=compression)
    r = read_geotiff(fp)
    assert r["array"].dtype == np.float32 and np.array_equal(r["array"], arr)
    assert r["transform"] == t and r["nodata"] == -9999.0 and (r["height"], r["width"]) == (300, 500)
    assert r["bounds"] == (1000.0, 5000.0, 2000.0, 5600.0)
    assert tuple(r["geo_tags"][TAG_GEOKEYS]) == UTM32


@pytest.mark.parametrize("compression", ["tiff_lzw", "raw", "tiff_adobe_deflate"])
def test_written_geotiff_is_read_by_an_independent_decoder(tmp_path, compression):
    """The files the run writes are what GDAL-side consumers read: OpenCV's bundled libtiff (not Pillow, which wrote them)
    must decode the same float32 samples.  A TIFF without georeferencing tags (what OpenCV writes) is refused on the way in."""
    cv2 = pytest.importorskip("cv2")
    arr = (np.random.default_rng(1).random((257, 333)) * 10 - 3).astype(np.float32)
    t = bounds_to_transform(0.0, 0.0, 333.0, 257.0, 333, 257)
    fp = write_geotiff(tmp_path / "w.tif", arr, t, nodata=-9999.0, geo_tags={TAG_GEOKEYS: UTM32}, compression=compression)
    seen = cv2.imread(str(fp), cv2.IMREAD_UNCHANGED)
    assert seen is not None and seen.dtype == np.float32 and np.array_equal(seen, arr)
    fp2 = tmp_path / "cv.tif"
    assert cv2.imwrite(str(fp2), arr)
    with pytest.raises(AssertionError, match="no georeferencing"):
        read_geotiff(fp2)


@pytest.mark.parametrize("nodata", [float("nan"), float("inf"), float("-inf"), -3.4028234663852886e38, 0.5])
def test_non_integer_and_non_finite_nodata_round_trips(tmp_path, nodata):
    """GDAL_NODATA "nan" is what float DEMs usually carry: writing the prediction next to such a DEM must not fail."""
    arr = np.ones((8, 8), np.float32)
    fp = write_geotiff(tmp_path / "n.tif", arr, (1.0, 0.0, 0.0, 0.0, -1.0, 8.0), nodata=nodata, geo_tags={TAG_GEOKEYS: UTM32})
    got = read_geotiff(fp)["nodata"]
    assert (np.isnan(got) and np.isnan(nodata)) or got == nodata


def test_pixel_is_point_shift_round_trips(tmp_path):
    arr = np.zeros((10, 20), np.float32)
    t = (2.0, 0.0, 100.0, 0.0, -2.0, 900.0)
    fp = write_geotiff(tmp_path / "p.tif", arr, t, geo_tags={TAG_GEOKEYS: UTM32_POINT})
    from PIL import Image
    with Image.open(fp) as img:
        assert tuple(img.tag_v2[33922])[3:5] == (101.0, 899.0)   # tie point at the centre of pixel (0, 0)
    assert read_geotiff(fp)["transform"] == t


def test_window_from_bounds_and_clip():
    t = (2.0, 0.0, 100.0, 0.0, -2.0, 900.0)
    assert window_from_bounds((100.0, 700.0, 300.0, 900.0), t) == (0, 0, 100, 100)
    assert window_from_bounds((110.0, 800.0, 151.0, 880.0), t) == (10, 5, 40, 21)       # lengths round up
    assert window_from_bounds((110.0004, 800.0, 150.0, 880.0), t) == (10, 5, 40, 20)    # sub-millipixel noise is snapped
    yy, xx = np.mgrid[0:100, 0:100]
    ras = {"array": (yy * 100 + xx).astype(np.float32), "transform": t, "height": 100, "width": 100}
    sub, ts = clip_to_bounds(ras, (110.0, 800.0, 150.0, 880.0))
    assert sub.shape == (40, 20) and sub[0, 0] == 10 * 100 + 5 and ts == (2.0, 0.0, 110.0, 0.0, -2.0, 880.0)
    assert bounds_of(ts, *sub.shape) == (110.0, 800.0, 150.0, 880.0)
    sub2, _ = clip_to_bounds(ras, (90.0, 650.0, 150.0, 880.0))  # window sticking out of the raster is intersected
    assert sub2.shape == (90, 25)
    with pytest.raises(AssertionError):
        clip_to_bounds(ras, (1000.0, 0.0, 1100.0, 10.0))


@pytest.mark.gpu
def test_tohr_files_matches_array_run(tmp_path, h1_model_fp):
    from floodsr_b200.raster_io import tohr_files
    from floodsr_b200.worker import ModelWorkerB200

    depth, dem_model_like = synth_raster(1024, 1024, seed=3)        # depth 64 x 64 cells of 16 m
    bounds = (300000.0, 5500000.0, 300000.0 + 1024.0, 5500000.0 + 1024.0)
    depth[5:8, 9:12] = -9999.0
    # DEM raster at 1.0 m on a grid that extends 40 m beyond the depth raster on every side
    big_bounds = (bounds[0] - 40.0, bounds[1] - 40.0, bounds[2] + 40.0, bounds[3] + 40.0)
    dem_big = np.pad(dem_model_like, 40, mode="edge").astype(np.float32)
    dem_big[600:610, 500:520] = -32768.0
    keys = {TAG_GEOKEYS: UTM32}
    depth_fp = write_geotiff(tmp_path / "depth.tif", depth, bounds_to_transform(*bounds, 64, 64), nodata=-9999.0, geo_tags=keys)
    dem_fp = write_geotiff(tmp_path / "dem.tif", dem_big, bounds_to_transform(*big_bounds, 1104, 1104), nodata=-32768.0, geo_tags=keys)
    res = tohr_files(depth_fp, dem_fp, tmp_path / "out" / "depth_hr.tif", h1_model_fp, precision="fp32")
    out = read_geotiff(res["output_fp"])
    assert (out["height"], out["width"]) == (1024, 1024) and out["bounds"] == bounds and out["nodata"] == -32768.0
    assert tuple(out["geo_tags"][TAG_GEOKEYS]) == UTM32
    assert np.array_equal(out["array"], res["prediction_m"])
    # the same run from arrays
    depth0 = np.where(depth == -9999.0, 0.0, depth).astype(np.float32)
    dem_crop = dem_big[40:-40, 40:-40].copy()
    dem_crop[dem_crop == -32768.0] = 0.0
    with ModelWorkerB200(h1_model_fp, precision="fp32") as worker:
        want = worker.run_raw_grids(depth0, bounds, dem_crop, bounds_to_transform(*bounds, 1024, 1024), dem_nodata=-32768.0)
    assert np.array_equal(res["prediction_m"], want["prediction_m"])
    assert not res["preprocess"]["post_resampled"] and res["preprocess"]["tile_cache_size"] == 9
    with pytest.raises(AssertionError, match="CRS mismatch"):
        other = write_geotiff(tmp_path / "d2.tif", depth, bounds_to_transform(*bounds, 64, 64), geo_tags={TAG_GEOKEYS: UTM32[:-1] + (32633,)})
        tohr_files(other, dem_fp, tmp_path / "x.tif", h1_model_fp)
