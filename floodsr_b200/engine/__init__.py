"""Engine package of the B200 backend; same export surface as the reference's `floodsr/engine/__init__.py`.

`EngineORT` is exported as an alias of `EngineB200` so that code written against the reference
(`from floodsr.engine import EngineORT`, `floodsr/models/ResUNet_16x_DEM.py:80,130`) runs on the B200
backend once this package is bound in its place (see `install_as_floodsr_engine` and INTEGRATION.md).
"""

from floodsr_b200.engine.b200 import EngineB200, ModelIOContract, install_as_floodsr_engine
from floodsr_b200.engine.base import EngineBase
from floodsr_b200.engine.providers import get_b200_info, get_onnxruntime_info, get_rasterio_info

EngineORT = EngineB200

__all__ = [
    "EngineB200",
    "EngineBase",
    "EngineORT",
    "ModelIOContract",
    "get_b200_info",
    "get_onnxruntime_info",
    "get_rasterio_info",
    "install_as_floodsr_engine",
]
