"""CPU oracle for the floodsr ToHR hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it, and only as the checker.  The product
package `floodsr_b200` never imports `oracle` (tests/test_host_logic.py::test_product_never_imports_the_oracle enforces this).

What restates what (reference paths relative to /root/reference):

* `oracle.preprocessing_np`  <- `floodsr/preprocessing.py:12-172`            (a5-a8, a11 of SURVEY.md section 8)
* `oracle.tiling_np`         <- `floodsr/tiling.py:7-45`                      (a12-a14)
* `oracle.stitch_np`         <- `floodsr/models/ResUNet_16x_DEM.py:140-393`   (a15, arrays instead of GeoTIFF paths)
* `oracle.engine_ref`        <- `floodsr/engine/ort.py:15-208`                (a2-a4, a9)
* `oracle.onnx_ref`          <- the ONNX Runtime session at `floodsr/engine/ort.py:54,193` (a10): the graph is
  executed operator by operator with torch CPU kernels in float32 (or float64 for error budgets).

Pinning status:

* preprocessing / tiling / tile-loop+stitch / run_tile orchestration: PINNED.  `tests/golden/make_golden.py`
  ran the reference's own `floodsr.preprocessing`, `floodsr.tiling`, `EngineORT.run_tile` and
  `ModelWorker._run_tiled_model_on_prepared` in the build container (only `onnxruntime.InferenceSession`
  and the GeoTIFF reader were stubbed) and committed the outputs under `tests/golden/`; the oracle must
  reproduce them bit for bit (`tests/test_oracle_golden.py`).
* network forward (a10): PARITY UNPINNED.  onnxruntime (pinned 1.24.2 / 1.22.2 by the reference's
  containers) and the `model_infer.onnx` release asset are absent offline and the reference holds no
  per-pixel golden output.  The interpreter is cross-checked against OpenCV's independent ONNX importer
  (`cv2.dnn.readNetFromONNX`) on the random-init H1 graph (`tests/test_forward_pin.py`), and
  `oracle.engine_ref.OracleEngine(backend="auto")` -- what every parity test and bench.py's CPU legs build --
  opens a live onnxruntime session (the reference's own call, ort.py:54,193) whenever onnxruntime is
  importable and reports which backend ran in `.backend`.
"""
