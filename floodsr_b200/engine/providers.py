"""Diagnostics with the reference's shape (`floodsr/engine/providers.py:6-29`) plus the B200 backend's own."""

from __future__ import annotations

import importlib.metadata as md


def get_onnxruntime_info() -> dict[str, object]:
    """ONNX Runtime is optional for this backend; report it without requiring it."""
    try:
        import onnxruntime as ort  # noqa: F401

        return {"installed": True, "version": md.version("onnxruntime"), "available_providers": list(ort.get_available_providers())}
    except Exception:
        return {"installed": False, "version": None, "available_providers": []}


def get_rasterio_info() -> dict[str, object]:
    try:
        return {"installed": True, "version": md.version("rasterio")}
    except md.PackageNotFoundError:
        return {"installed": False, "version": None}


def get_b200_info() -> dict[str, object]:
    """Native library + device diagnostics of the B200 backend."""
    from floodsr_b200 import _lib

    info: dict[str, object] = {"installed": _lib.LIB_PATH.exists(), "library": str(_lib.LIB_PATH), "devices": 0, "abi": None}
    if info["installed"]:
        lib = _lib.load_library()
        info["abi"] = int(lib.fsr_abi_version())
        info["devices"] = int(lib.fsr_device_count())
    return info
