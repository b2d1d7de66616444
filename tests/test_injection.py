"""The zero-edit route into the reference: `install_as_floodsr_engine()` makes the unmodified worker module build an
`EngineB200` where it says `EngineORT(...)` (`floodsr/models/ResUNet_16x_DEM.py:80,130`, `floodsr/model_registry.py:373-394`).

CPU part: needs the reference tree (skipped where /root/reference is absent, e.g. on the GPU box).  The GPU part drives
the reference worker's tile loop, restated in oracle/stitch_np.py, through `EngineB200.run_tile` tile by tile.
"""

from __future__ import annotations

import ast
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")

_CHILD = r"""
import sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, {ref!r})
from floodsr_b200.engine import EngineB200, install_as_floodsr_engine
install_as_floodsr_engine()
import floodsr.engine as fe
from floodsr.model_registry import resolve_model_worker_class
worker_cls = resolve_model_worker_class("ResUNet_16x_DEM")
g = worker_cls.__enter__.__globals__
assert g["EngineORT"] is EngineB200, g["EngineORT"]
assert fe.EngineORT is EngineB200
print("EXPORTS", sorted(n for n in ("EngineORT", "get_onnxruntime_info", "get_rasterio_info") if hasattr(fe, n)))
info = fe.get_onnxruntime_info()
print("INFO", sorted(info))
import inspect
print("SIG", list(inspect.signature(EngineB200.__init__).parameters)[:4], list(inspect.signature(EngineB200.run_tile).parameters))
"""


@pytest.mark.skipif(not (REFERENCE / "floodsr" / "engine" / "__init__.py").exists(), reason="reference tree not present")
def test_install_routes_the_unmodified_reference_worker_to_engine_b200():
    # a child process: the shim registers modules under the reference's names, which must not leak into this test session
    out = subprocess.run([sys.executable, "-c", _CHILD.format(repo=str(REPO), ref=str(REFERENCE))], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = dict(l.split(" ", 1) for l in out.stdout.splitlines() if " " in l)
    # the shim exports what floodsr/engine/__init__.py:3-8 exports
    tree = ast.parse((REFERENCE / "floodsr" / "engine" / "__init__.py").read_text())
    ref_all = next(ast.literal_eval(n.value) for n in tree.body if isinstance(n, ast.Assign) and n.targets[0].id == "__all__")
    assert ast.literal_eval(lines["EXPORTS"]) == sorted(ref_all)
    # constructor and run_tile keep EngineORT's parameter names (ort.py:31-36,128-138)
    ref_ort = ast.parse((REFERENCE / "floodsr" / "engine" / "ort.py").read_text())
    cls = next(n for n in ref_ort.body if isinstance(n, ast.ClassDef) and n.name == "EngineORT")
    fn = {f.name: [a.arg for a in f.args.args] for f in cls.body if isinstance(f, ast.FunctionDef)}
    ctor, run_tile = ast.literal_eval(lines["SIG"].split("] [")[0] + "]"), ast.literal_eval("[" + lines["SIG"].split("] [")[1])
    assert ctor == fn["__init__"][:4]
    assert run_tile == fn["run_tile"]


@pytest.mark.gpu
def test_reference_tile_loop_over_engine_b200_equals_run_raster(h1_model_fp):
    """The reference worker calls `engine.run_tile` once per window and blends on the host
    (`ResUNet_16x_DEM.py:250-363`); `run_raster` does the whole loop on the device.  Same tiles, same window order:
    the two mosaics must be identical."""
    from floodsr_b200.engine import EngineB200
    from floodsr_b200.synth import synth_raster
    from oracle.stitch_np import run_tiled

    eng = EngineB200(h1_model_fp)  # default precision: the fp32-tolerance mode
    for (h, w, method) in ((1024, 1536, "feather"), (976, 1104, "feather"), (1024, 1024, "hard")):
        depth, dem = synth_raster(h, w, seed=h + 3 * w)
        want, n_tiles, summary = run_tiled(eng, depth, dem, window_method=method, overlap_lr=8)
        got, got_n, got_summary = eng.run_raster(depth, dem, window_method=method, overlap_lr=8)
        assert got_n == n_tiles and got_summary == summary
        assert np.array_equal(got, want), (h, w, method, float(np.abs(got - want).max()))
    eng.close()
