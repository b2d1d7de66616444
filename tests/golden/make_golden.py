"""Generate tests/golden/*.json|npz by running the REFERENCE's own code in the build container.

Run once from the repo root (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference: `floodsr.preprocessing` (a5-a8, a11), `floodsr.tiling`
(a12-a14), `floodsr.engine.ort.EngineORT` (a2-a4, a9) and
`floodsr.models.ResUNet_16x_DEM.ModelWorker._run_tiled_model_on_prepared` (a15).
What is stubbed, because it is absent offline: the `onnxruntime` module (its `InferenceSession` is
replaced by (a) an exactly-reproducible analytic function or (b) the torch oracle interpreter on the
random-init H1 graph) and the rasterio GeoTIFF reader (`_read_single_band_raster` returns in-memory
arrays).  Fixtures hold sha256 digests of the reference outputs plus strided samples, so they stay small.
"""

from __future__ import annotations

import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

from floodsr_b200.synth import synth_dem, synth_depth, synth_raster, synth_tile  # noqa: E402


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    # -0.0 and +0.0 compare equal in the reference's own tests (np.array_equal); canonicalise them
    if a.dtype.kind == "f":
        a = a + np.zeros((), dtype=a.dtype)
    return hashlib.sha256(a.tobytes()).hexdigest()


def analytic_forward(depth_nhwc: np.ndarray, dem_nhwc: np.ndarray) -> np.ndarray:
    """Bit-reproducible stand-in network: 0.5*nearest16(depth) + 0.25*dem - 1/16 (float32)."""
    d = depth_nhwc.astype(np.float32)
    s = dem_nhwc.shape[1] // d.shape[1]
    up = np.repeat(np.repeat(d, s, axis=1), s, axis=2)
    return (np.float32(0.5) * up + np.float32(0.25) * dem_nhwc.astype(np.float32)) - np.float32(0.0625)


class _Meta:
    def __init__(self, name, shape):
        self.name, self.shape = name, shape


def install_onnxruntime_stub(forward):
    mod = types.ModuleType("onnxruntime")

    class InferenceSession:
        def __init__(self, path, providers=None, **kw):
            self._providers = list(providers or ["CPUExecutionProvider"])

        def get_inputs(self):
            return [_Meta("depth_lr", ["unk__300", 32, 32, 1]), _Meta("dem_hr", ["unk__301", 512, 512, 1])]

        def get_outputs(self):
            return [_Meta("depth_hr_pred", ["unk__302", 512, 512, 1])]

        def get_providers(self):
            return self._providers

        def run(self, output_names, feed):
            return [forward(feed["depth_lr"], feed["dem_hr"])]

    mod.InferenceSession = InferenceSession
    mod.get_available_providers = lambda: ["CPUExecutionProvider"]
    sys.modules["onnxruntime"] = mod
    return mod


def tile_cases():
    """(name, depth, dem, kwargs) single-tile inputs."""
    cases = []
    d0, e0 = synth_tile(0)
    cases.append(("synth0", d0, e0, {}))
    d1, e1 = synth_tile(1)
    cases.append(("synth1_pct50", d1, e1, {"dem_pct_clip": 50.0}))
    cases.append(("synth1_pct100", d1, e1, {"dem_pct_clip": 100.0}))
    cases.append(("synth1_pct99p9_depth3", d1, e1, {"dem_pct_clip": 99.9, "max_depth": 3.0}))
    # the reference's own contract fixture (tests/conftest.py:148-156)
    cases.append(
        (
            "ort_tile_inputs",
            np.full((32, 32), 1.5, dtype=np.float32),
            np.linspace(500.0, 1000.0, 512 * 512, dtype=np.float32).reshape((512, 512)),
            {"depth_lr_nodata": -9999.0, "dem_hr_nodata": -9999.0},
        )
    )
    # nodata cells + zero-padded edge (as produced by the worker's np.pad) + negative elevations
    d2, e2 = synth_tile(2)
    e2 = e2.copy()
    e2[300:, :] = 0.0
    e2[:, 400:] = 0.0
    e2[10:20, 10:40] = -9999.0
    e2[50:60, 50:60] = -3.5
    d2 = d2.copy()
    d2[3:5, 7:9] = -9999.0
    cases.append(("padded_nodata", d2, e2, {"depth_lr_nodata": -9999.0, "dem_hr_nodata": -9999.0}))
    # narrow-range DEM (all values share their high float bits) and many exact ties
    rng = np.random.default_rng(7)
    e3 = (1073.0 + np.round(rng.random((512, 512)) * 17.0, 1)).astype(np.float32)
    cases.append(("narrow_ties", synth_depth(32, 32, 3), e3, {}))
    # all-zero DEM tile: reference returns zeros without raising (preprocessing.py:72-80)
    cases.append(("zero_dem", synth_depth(32, 32, 4), np.zeros((512, 512), dtype=np.float32), {}))
    return cases


def raster_cases():
    """(name, depth_lr, dem_hr, kwargs) model-space rasters for the tile loop."""
    cases = []
    d, e = synth_raster(1024, 1536, seed=11)
    cases.append(("r1024x1536_feather", d, e, {"window_method": "feather", "overlap_lr": 8}))
    cases.append(("r1024x1536_hard", d, e, {"window_method": "hard", "overlap_lr": 8}))
    d, e = synth_raster(976, 1104, seed=12)
    cases.append(("r976x1104_feather_pad", d, e, {"window_method": "feather", "overlap_lr": 8}))
    cases.append(("r976x1104_feather_ov4", d, e, {"window_method": "feather", "overlap_lr": 4}))
    d, e = synth_raster(528, 1296, seed=13)  # trailing-edge start 1 px... several tiles deep in x
    cases.append(("r528x1296_feather_ov12", d, e, {"window_method": "feather", "overlap_lr": 12}))
    d, e = synth_raster(512, 512, seed=14)
    cases.append(("r512_single", d, e, {"window_method": "feather", "overlap_lr": 8}))
    return cases


def sample(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a[::37, ::41]).copy()


def main():
    assert REF.exists(), "needs /root/reference"
    sys.path.insert(0, str(REF))
    meta: dict = {"numpy": np.__version__, "cases": {}}
    arrays: dict[str, np.ndarray] = {}

    # ---- a5-a8, a11: floodsr.preprocessing directly -------------------------------------------
    import floodsr.preprocessing as rp
    import floodsr.tiling as rt

    for name, depth, dem, kw in tile_cases():
        max_depth = kw.get("max_depth", 5.0)
        pct = kw.get("dem_pct_clip", 95.0)
        dz = rp.replace_nodata_with_zero(depth, kw.get("depth_lr_nodata"))
        ez = rp.replace_nodata_with_zero(dem, kw.get("dem_hr_nodata"))
        dn = rp.scale_depth_log1p_np(dz, max_depth=max_depth)
        en, stats = rp.normalize_dem(ez, pct_clip=pct)
        inv = rp.invert_depth_log1p_np(dn, max_depth=max_depth)
        meta["cases"][f"pre/{name}"] = {
            "kwargs": kw,
            "stats": {k: float(v).hex() for k, v in stats.items()},
            "depth_norm_sha": digest(dn),
            "dem_norm_sha": digest(en),
            "invert_sha": digest(inv),
        }
        arrays[f"pre/{name}/dem_norm_s"] = sample(en)
        arrays[f"pre/{name}/depth_norm"] = dn
    # flat non-zero DEM must raise (preprocessing.py:82)
    try:
        rp.normalize_dem(np.full((512, 512), 7.0, dtype=np.float32))
        meta["flat_nonzero_raises"] = False
    except AssertionError as exc:
        meta["flat_nonzero_raises"] = str(exc)

    # ---- a12-a14: floodsr.tiling ---------------------------------------------------------------
    starts = {}
    for total in [512, 513, 640, 895, 896, 897, 1024, 1296, 1536, 4096, 8192, 32768]:
        for tile, stride in [(512, 384), (512, 448), (512, 320), (512, 512), (512, 1)]:
            if stride == 1 and total > 640:
                continue
            starts[f"{total},{tile},{stride}"] = rt.build_tile_starts(total, tile, stride)
    meta["tile_starts"] = starts
    for tile, ov in [(512, 128), (512, 64), (512, 192), (512, 0), (512, 511), (64, 16)]:
        arrays[f"ramp/{tile},{ov}"] = rt.build_feather_ramp(tile, ov)
    meta["origins_3x2"] = [list(t) for t in rt.iter_window_origins([0, 384, 512], [0, 100], use_progress=False)]

    # ---- a2-a4, a9, a15 with the analytic network ----------------------------------------------
    install_onnxruntime_stub(analytic_forward)
    from floodsr.engine.ort import EngineORT
    import importlib.util

    model_stub = REF / "tests" / "data" / "model_infer_dummy.onnx"  # any existing file; the stub ignores it
    eng = EngineORT(model_stub)
    meta["contract"] = {
        "depth_lr_hwc": list(eng.contract.depth_lr_hwc),
        "dem_hr_hwc": list(eng.contract.dem_hr_hwc),
        "output_hwc": list(eng.contract.output_hwc),
        "scale": eng.contract.scale,
        "output_name": eng.contract.output_name,
    }
    for name, depth, dem, kw in tile_cases():
        res = eng.run_tile(depth, dem, **kw)
        meta["cases"][f"run_tile_analytic/{name}"] = {
            "prediction_m_sha": digest(res["prediction_m"]),
            "prediction_norm_sha": digest(res["prediction_norm"]),
            "stats": {k: float(v).hex() for k, v in res["dem_stats_used"].items()},
        }
        arrays[f"run_tile_analytic/{name}/pred_m_s"] = sample(res["prediction_m"])
    # normalize_inputs=False branch (ort.py:163-180)
    dn = rp.scale_depth_log1p_np(tile_cases()[0][1], 5.0)
    en, _ = rp.normalize_dem(tile_cases()[0][2])
    res = eng.run_tile(dn, en, normalize_inputs=False)
    meta["cases"]["run_tile_analytic/prenormalised"] = {
        "prediction_m_sha": digest(res["prediction_m"]),
        "stats": {k: float(v).hex() for k, v in res["dem_stats_used"].items()},
    }

    spec = importlib.util.spec_from_file_location("ref_worker", REF / "floodsr" / "models" / "ResUNet_16x_DEM.py")
    wm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wm)
    store: dict[str, np.ndarray] = {}
    wm._read_single_band_raster = lambda fp: (store[str(fp)], None, {})
    for name, depth, dem, kw in raster_cases():
        store["depth"], store["dem"] = depth, dem
        with wm.ModelWorker(model_stub) as worker:
            out, n_tiles, summary = worker._run_tiled_model_on_prepared(
                depth_lr_fp="depth",
                dem_hr_fp="dem",
                preprocess_cfg={"max_depth": 5.0, "dem_pct_clip": 95.0},
                model_lr_tile=32,
                model_scale=16,
                contract_hr_tile=512,
                **kw,
            )
        meta["cases"][f"raster_analytic/{name}"] = {
            "kwargs": kw,
            "shape": list(out.shape),
            "n_tiles": int(n_tiles),
            "sha": digest(out),
            "summary": {k: float(v).hex() for k, v in summary.items()},
        }
        arrays[f"raster_analytic/{name}/s"] = sample(out)

    # ---- a10 via the torch interpreter on the random-init H1 graph (tolerance-level fixture) ----
    from floodsr_b200.h1 import write_h1_model
    from oracle.onnx_ref import RefSession
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        model_fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
        sess = RefSession(model_fp)
        install_onnxruntime_stub(lambda d, e: sess.run({"depth_lr": d, "dem_hr": e})[0])
        import importlib

        import floodsr.engine.ort as ort_mod

        importlib.reload(ort_mod)
        eng = ort_mod.EngineORT(model_fp)
        for name, depth, dem, kw in tile_cases()[:1] + tile_cases()[4:6]:
            res = eng.run_tile(depth, dem, **kw)
            arrays[f"run_tile_h1/{name}/pred_m_s"] = sample(res["prediction_m"])
            arrays[f"run_tile_h1/{name}/pred_norm_s"] = sample(res["prediction_norm"])
            meta["cases"][f"run_tile_h1/{name}"] = {
                "pred_norm_min": float(res["prediction_norm"].min()),
                "pred_norm_max": float(res["prediction_norm"].max()),
                "pred_m_mean": float(res["prediction_m"].mean()),
            }

    (OUT / "golden.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
    np.savez_compressed(OUT / "golden.npz", **arrays)
    print("wrote", OUT / "golden.json", OUT / "golden.npz", sum(a.nbytes for a in arrays.values()), "bytes of arrays")


if __name__ == "__main__":
    main()
