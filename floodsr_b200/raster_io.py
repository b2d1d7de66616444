"""Single-band GeoTIFF in / out without rasterio, and a file-to-file ToHR run without temp rasters (SURVEY.md section 8 f#3).

The reference reads its two inputs with rasterio (`floodsr/preprocessing.py:304-360`), writes the aligned rasters to a
temporary directory as LZW GeoTIFFs (`write_prepared_rasters`, `:411-473`), re-reads them inside the worker
(`ResUNet_16x_DEM.py:184-185`) and writes the result with the DEM's profile (`:586-604`).  Here the decoded rasters go
straight into (optionally page-locked) host buffers that `EngineB200` copies from, the aligned rasters never touch the
disk, and the GeoTIFF codec is Pillow/libtiff:

    read_geotiff     float32 array + affine transform + nodata + the georeferencing tags, verbatim
    window_from_bounds / clip_to_bounds   `rasterio.windows.from_bounds(...).round_offsets().round_lengths()` (`:352`)
    write_geotiff    float32, LZW by default, georeferencing tags carried over from the DEM
    tohr_files       depth.tif + dem.tif -> depth_hr.tif through `ModelWorkerB200.run_raw_grids`

Only north-up rasters (ModelPixelScale + ModelTiepoint) are handled, which is what the reference accepts after its CRS and
bounds checks; CRS equality is checked on the raw GeoKey directory instead of through PROJ.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

TAG_PIXEL_SCALE = 33550      # ModelPixelScaleTag   (sx, sy, sz)
TAG_TIEPOINT = 33922         # ModelTiepointTag     (i, j, k, x, y, z)
TAG_TRANSFORMATION = 34264   # ModelTransformationTag (4 x 4, row-major)
TAG_GEOKEYS = 34735          # GeoKeyDirectoryTag
TAG_GEO_DOUBLES = 34736
TAG_GEO_ASCII = 34737
TAG_GDAL_NODATA = 42113      # ASCII
GEO_TAGS = (TAG_PIXEL_SCALE, TAG_TIEPOINT, TAG_TRANSFORMATION, TAG_GEOKEYS, TAG_GEO_DOUBLES, TAG_GEO_ASCII)
KEY_RASTER_TYPE = 1025       # GTRasterTypeGeoKey: 1 = PixelIsArea, 2 = PixelIsPoint


def _pil():
    from PIL import Image, TiffImagePlugin, TiffTags  # noqa: PLC0415

    Image.MAX_IMAGE_PIXELS = None  # 32k x 32k DEMs are legitimate here
    return Image, TiffImagePlugin, TiffTags


def _raster_type(geokeys) -> int:
    if not geokeys or len(geokeys) < 4:
        return 1
    for k in range(int(geokeys[3])):
        key, loc, _count, value = geokeys[4 + 4 * k: 8 + 4 * k]
        if key == KEY_RASTER_TYPE and loc == 0:
            return int(value)
    return 1


def read_geotiff(fp, pinned: bool = False) -> dict:
    """Decode band 1 as float32 (`ds.read(1).astype(float32)`, preprocessing.py:345) with its georeferencing.

    `pinned=True` returns the pixels in page-locked memory (`fsr_host_alloc`) so that `EngineB200.run_raster` overlaps the
    H2D copies with the kernels.
    """
    Image, _, _ = _pil()
    path = Path(fp).expanduser().resolve()
    assert path.exists(), f"raster does not exist: {path}"
    with Image.open(path) as img:
        assert getattr(img, "n_frames", 1) >= 1
        bands = img.getbands()
        assert len(bands) == 1, f"raster must have 1 band; got {len(bands)}"
        tags = {t: img.tag_v2.get(t) for t in GEO_TAGS if img.tag_v2.get(t) is not None}
        nodata_raw = img.tag_v2.get(TAG_GDAL_NODATA)
        decoded = np.asarray(img)
    if pinned:
        from floodsr_b200 import _lib  # noqa: PLC0415

        array = _lib.pinned_empty(decoded.shape)
        np.copyto(array, decoded, casting="unsafe")
    else:
        array = np.ascontiguousarray(decoded, dtype=np.float32)
    nodata = None
    if nodata_raw is not None:
        text = (nodata_raw if isinstance(nodata_raw, str) else str(nodata_raw[0])).strip().strip("\x00")
        nodata = float(text) if text else None
    transform = _transform_from_tags(tags)
    h, w = array.shape
    return {"array": array, "transform": transform, "nodata": nodata, "height": h, "width": w, "geo_tags": tags,
            "bounds": bounds_of(transform, h, w), "path": path}


def _transform_from_tags(tags: dict):
    if TAG_PIXEL_SCALE in tags and TAG_TIEPOINT in tags:
        sx, sy = float(tags[TAG_PIXEL_SCALE][0]), float(tags[TAG_PIXEL_SCALE][1])
        i, j, _k, x, y, _z = (float(v) for v in tags[TAG_TIEPOINT][:6])
        c, f = x - i * sx, y + j * sy
        if _raster_type(tags.get(TAG_GEOKEYS)) == 2:  # PixelIsPoint: the tie point is a pixel centre
            c, f = c - 0.5 * sx, f + 0.5 * sy
        return (sx, 0.0, c, 0.0, -sy, f)
    if TAG_TRANSFORMATION in tags:
        m = [float(v) for v in tags[TAG_TRANSFORMATION]]
        assert m[1] == 0.0 and m[4] == 0.0, "rotated rasters are not supported"
        return (m[0], 0.0, m[3], 0.0, m[5], m[7])
    raise AssertionError("raster has no georeferencing (ModelPixelScale + ModelTiepoint)")


def bounds_of(transform, height: int, width: int):
    """(west, south, east, north) like `rasterio.DatasetReader.bounds`."""
    a, _, c, _, e, f = transform
    return (c, f + e * height, c + a * width, f)


def window_from_bounds(bounds, transform, pixel_precision: int = 3):
    """`from_bounds(*bounds, transform).round_offsets().round_lengths()` (preprocessing.py:352): (row_off, col_off, h, w).

    rasterio rounds the fractional offsets down and the fractional lengths up after snapping both to `pixel_precision`
    decimals; restated here (rasterio is not importable offline, so this rule is unpinned against it).
    """
    west, south, east, north = (float(v) for v in bounds)
    a, _, c, _, e, f = transform
    col0, col1 = (west - c) / a, (east - c) / a
    row0, row1 = (north - f) / e, (south - f) / e
    col_off = math.floor(round(min(col0, col1), pixel_precision))
    row_off = math.floor(round(min(row0, row1), pixel_precision))
    width = math.ceil(round(abs(col1 - col0), pixel_precision))
    height = math.ceil(round(abs(row1 - row0), pixel_precision))
    return int(row_off), int(col_off), int(height), int(width)


def clip_to_bounds(raster: dict, bounds):
    """The DEM pixels covering `bounds` on the DEM's own grid and their transform (`ds.read(1, window=...)`, `window_transform`).

    Like rasterio's windowed read without `boundless`, the window is intersected with the raster.
    """
    row_off, col_off, h, w = window_from_bounds(bounds, raster["transform"])
    r0, c0 = max(row_off, 0), max(col_off, 0)
    r1, c1 = min(row_off + h, raster["height"]), min(col_off + w, raster["width"])
    assert r1 > r0 and c1 > c0, f"clipped DEM is empty for bounds {tuple(bounds)}"
    a, _, c, _, e, f = raster["transform"]
    return raster["array"][r0:r1, c0:c1], (a, 0.0, c + a * c0, 0.0, e, f + e * r0)


def write_geotiff(fp, array: np.ndarray, transform, nodata=None, geo_tags: dict | None = None, compression: str = "tiff_lzw") -> Path:
    """float32 single-band GeoTIFF with the given transform; CRS keys are carried over from `geo_tags` (the DEM's)."""
    Image, TiffImagePlugin, TiffTags = _pil()
    path = Path(fp).expanduser()
    path.parent.mkdir(parents=True, exist_ok=True)
    arr = np.ascontiguousarray(array, dtype=np.float32)
    assert arr.ndim == 2, f"expected a 2-D raster; got {arr.shape}"
    a, b, c, d, e, f = (float(v) for v in tuple(transform)[:6])
    assert b == 0.0 and d == 0.0, "rotated rasters are not supported"
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[TAG_PIXEL_SCALE] = (a, -e, 0.0)
    ifd.tagtype[TAG_PIXEL_SCALE] = TiffTags.DOUBLE
    # the tie point is written for the outer corner of pixel (0, 0); PixelIsPoint rasters get the half-pixel shift back
    point = _raster_type((geo_tags or {}).get(TAG_GEOKEYS)) == 2
    ifd[TAG_TIEPOINT] = (0.0, 0.0, 0.0, c + (0.5 * a if point else 0.0), f + (0.5 * e if point else 0.0), 0.0)
    ifd.tagtype[TAG_TIEPOINT] = TiffTags.DOUBLE
    for tag, kind in ((TAG_GEOKEYS, TiffTags.SHORT), (TAG_GEO_DOUBLES, TiffTags.DOUBLE), (TAG_GEO_ASCII, TiffTags.ASCII)):
        if geo_tags and geo_tags.get(tag) is not None:
            ifd[tag] = geo_tags[tag]
            ifd.tagtype[tag] = kind
    if nodata is not None:
        v = float(nodata)  # GDAL_NODATA is ASCII: "nan", "inf", "-inf" are legal and common for float DEMs
        ifd[TAG_GDAL_NODATA] = ("nan" if math.isnan(v) else ("inf" if v > 0 else "-inf")) if not math.isfinite(v) else (
            str(int(v)) if v.is_integer() else repr(v))
        ifd.tagtype[TAG_GDAL_NODATA] = TiffTags.ASCII
    Image.fromarray(arr, mode="F").save(path, format="TIFF", compression=compression, tiffinfo=ifd)
    return path


def tohr_files(depth_lr_fp, dem_hr_fp, output_fp, model_fp, *, max_depth: float | None = None, dem_pct_clip: float | None = None,
               window_method: str = "feather", tile_overlap: int | None = None, tile_size: int | None = None,
               precision: str | None = None, logger=None) -> dict:
    """`ModelWorker.run` (ResUNet_16x_DEM.py:395-640) from files to a file on the B200 engine, no temporary rasters."""
    from floodsr_b200.worker import ModelWorkerB200  # noqa: PLC0415

    depth = read_geotiff(depth_lr_fp)
    dem = read_geotiff(dem_hr_fp)
    # CRS agreement (preprocessing.py:308-323) on the raw GeoKey directories; a depth raster without keys takes the DEM's
    dk, mk = depth["geo_tags"].get(TAG_GEOKEYS), dem["geo_tags"].get(TAG_GEOKEYS)
    assert mk is not None, "both rasters must define CRS"
    if dk is not None:
        assert tuple(dk) == tuple(mk), f"CRS mismatch\n    depth={dk}\n    dem={mk}"
    depth_lr = depth["array"]
    if depth["nodata"] is not None:  # replace_nodata_with_zero (preprocessing.py:167-172)
        depth_lr = np.where(np.isclose(depth_lr, depth["nodata"]), 0.0, depth_lr).astype(np.float32, copy=False)
    dem_crop, dem_crop_transform = clip_to_bounds(dem, depth["bounds"])
    if dem["nodata"] is not None:
        dem_crop_clean = np.where(np.isclose(dem_crop, dem["nodata"]), 0.0, dem_crop).astype(np.float32, copy=False)
    else:
        dem_crop_clean = np.ascontiguousarray(dem_crop, dtype=np.float32)
    if not np.isfinite(dem_crop_clean).all():
        raise AssertionError("DEM contains non-finite values after clipping")
    with ModelWorkerB200(model_fp, logger=logger, precision=precision) as worker:
        # the reference resamples the nodata-replaced crop but still passes the nodata value to reproject (:345-387)
        res = worker.run_raw_grids(depth_lr, depth["bounds"], dem_crop_clean, dem_crop_transform, dem_nodata=dem["nodata"],
                                   max_depth=max_depth, dem_pct_clip=dem_pct_clip, window_method=window_method,
                                   tile_overlap=tile_overlap, tile_size=tile_size)
    # output on the clipped DEM grid with the DEM's profile (ResUNet_16x_DEM.py:548-551, :589), then the reference's
    # read-back checks of shape and bounds (:590-598)
    out_bounds = bounds_of(dem_crop_transform, *res["prediction_m"].shape)
    assert all(np.isclose(x, y, atol=1e-6, rtol=0.0) for x, y in zip(out_bounds, depth["bounds"])), (
        f"output profile bounds {out_bounds} do not match incoming low-res bounds {depth['bounds']}"
    )
    out_path = write_geotiff(output_fp, res["prediction_m"], dem_crop_transform, nodata=dem["nodata"], geo_tags=dem["geo_tags"])
    written = read_geotiff(out_path)
    assert (written["height"], written["width"]) == tuple(dem_crop.shape), (
        f"written output shape {(written['height'], written['width'])} must match raw DEM shape {tuple(dem_crop.shape)}"
    )
    assert all(np.isclose(x, y, atol=1e-6, rtol=0.0) for x, y in zip(written["bounds"], depth["bounds"])), (
        f"written output bounds {written['bounds']} must match incoming low-res bounds {depth['bounds']}"
    )
    res["output_fp"] = str(out_path)
    res["output_transform"] = dem_crop_transform
    return res
