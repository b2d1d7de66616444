"""Shared pytest configuration: the `gpu` marker, repo-root import path, session fixtures."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Reference-generated fixtures (tests/golden/make_golden.py)."""
    meta = json.loads((REPO / "tests" / "golden" / "golden.json").read_text())
    arrays = np.load(REPO / "tests" / "golden" / "golden.npz")
    return meta, arrays


@pytest.fixture(scope="session")
def h1_model_fp(tmp_path_factory):
    """Random-init H1 `model_infer.onnx` (seed 0), same lookup role as the reference's `tohr_model_fp`."""
    from floodsr_b200.h1 import write_h1_model

    return write_h1_model(tmp_path_factory.mktemp("ResUNet_16x_DEM") / "model_infer.onnx", seed=0)


@pytest.fixture(scope="session")
def ort_tile_inputs():
    """The reference's synthetic contract tile (/root/reference/tests/conftest.py:148-156)."""
    return {
        "depth_lr": np.full((32, 32), 1.5, dtype=np.float32),
        "dem_hr": np.linspace(500.0, 1000.0, 512 * 512, dtype=np.float32).reshape((512, 512)),
        "depth_lr_nodata": -9999.0,
        "dem_hr_nodata": -9999.0,
    }
