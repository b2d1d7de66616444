"""Independent pins of the network forward (a10), the one stage whose arithmetic lives in third-party onnxruntime
(`floodsr/engine/ort.py:54,193`; pinned 1.24.2 / 1.22.2 by the reference, absent offline).

* OpenCV's ONNX importer (`cv2.dnn.readNetFromONNX`) is a second, unrelated implementation of the same operators: it must
  agree with the oracle's torch interpreter on the same file.
* When onnxruntime is importable, the reference's own call (`InferenceSession(...).run`, CPUExecutionProvider) is the
  oracle of record: the torch interpreter must agree with it, and the GPU parity tests compare against it first.
"""

from __future__ import annotations

import numpy as np
import pytest

from floodsr_b200.synth import synth_tile
from oracle import preprocessing_np as pp
from oracle.engine_ref import OracleEngine


def _norm_inputs(seed):
    depth, dem = synth_tile(seed)
    return pp.scale_depth_log1p(depth, 5.0)[None], pp.normalize_dem(dem)[0][None]


def test_opencv_onnx_importer_agrees_with_the_oracle_interpreter(h1_model_fp):
    cv2 = pytest.importorskip("cv2")
    net = cv2.dnn.readNetFromONNX(str(h1_model_fp))
    assert len(net.getLayerNames()) > 40  # the whole graph was imported, not a stub
    eng = OracleEngine(h1_model_fp)
    for seed in (3, 11):
        dn, en = _norm_inputs(seed)
        net.setInput(np.ascontiguousarray(dn[..., None]), "depth_lr")
        net.setInput(np.ascontiguousarray(en[..., None]), "dem_hr")
        got = net.forward().reshape(1, 512, 512)
        want = eng.forward_norm(dn, en)
        assert want.std() > 0.02
        assert np.abs(got - want).max() <= 1e-5  # two fp32 implementations with different summation orders


def test_live_onnxruntime_session_agrees_with_the_oracle_interpreter(h1_model_fp):
    ort = pytest.importorskip("onnxruntime")  # not installed in the offline image: skipped there, reported as such
    sess = ort.InferenceSession(str(h1_model_fp), providers=["CPUExecutionProvider"])  # ort.py:54
    eng = OracleEngine(h1_model_fp)
    for seed in (3, 11):
        dn, en = _norm_inputs(seed)
        got = sess.run([eng.output_name], {"depth_lr": dn[..., None], "dem_hr": en[..., None]})[0][..., 0]  # ort.py:193
        assert np.abs(got - eng.forward_norm(dn, en)).max() <= 1e-5
