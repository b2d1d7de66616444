"""Engine contract, mirror of the reference's `floodsr/engine/base.py:10-29`."""

from __future__ import annotations

from abc import ABC, abstractmethod
from pathlib import Path
from typing import Any

import numpy as np


class EngineBase(ABC):
    """Abstract inference engine: `load()`, `run_tile(depth_lr_m, dem_hr_m, **kwargs) -> dict`, `model_path()`."""

    @abstractmethod
    def load(self) -> None:
        """Load model resources."""

    @abstractmethod
    def run_tile(self, depth_lr_m: np.ndarray, dem_hr_m: np.ndarray, **kwargs: Any) -> dict[str, Any]:
        """One inference pass for a low-res depth tile + high-res DEM tile."""

    @abstractmethod
    def model_path(self) -> Path:
        """Path of the model artefact this engine was built from."""
