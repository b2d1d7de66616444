"""Locating and verifying the model weights the way the reference does; no download.

Mirrors the lookup side of `floodsr/model_registry.py:309-336` with `floodsr/cache_paths.py:13-40` and
`floodsr/checksums.py:11-45`: the weights of version `V` live at `<user cache dir of app "floodsr">/<V>/<file_name>` (on Linux
`~/.cache/floodsr/ResUNet_16x_DEM/model_infer.onnx`), the manifest (`floodsr/models.json`) names file and sha256, and tests
look in `./_inputs/<V>/*.onnx` first (`tests/conftest.py:200-215`).  Fetching over the network stays in the reference CLI
(`floodsr models fetch`); without a cached asset the engine can run on the random-init H1 graph (`floodsr_b200/h1.py`).
"""

from __future__ import annotations

import hashlib
import json
import os
import sys
from pathlib import Path

APP_NAME = "floodsr"
# manifest entry of the reference release (floodsr/models.json:3-8)
DEFAULT_MANIFEST = {
    "models": {
        "ResUNet_16x_DEM": {
            "file_name": "model_infer.onnx",
            "url": "https://github.com/cefect/floodsr/releases/download/v2026.02.19/model_infer.onnx",
            "sha256": "ea907c0f3bd482f436b26b4d28df64a471af51e0779203a50125e773eeda9c4a",
            "description": "16x DEM-conditioned ResUNet",
        }
    }
}


def user_cache_dir() -> Path:
    """`platformdirs.user_cache_dir("floodsr", "floodsr")` (the call at `cache_paths.py:21`), without needing the package."""
    try:
        from platformdirs import user_cache_dir as _ucd  # the reference's own dependency, when installed

        return Path(_ucd(APP_NAME, APP_NAME))
    except ImportError:
        if sys.platform == "darwin":
            return Path.home() / "Library" / "Caches" / APP_NAME
        if os.name == "nt":
            base = os.environ.get("LOCALAPPDATA") or str(Path.home() / "AppData" / "Local")
            return Path(base) / APP_NAME / APP_NAME / "Cache"
        return Path(os.environ.get("XDG_CACHE_HOME") or (Path.home() / ".cache")) / APP_NAME


def model_cache_path(model_version: str, file_name: str, cache_dir: str | Path | None = None) -> Path:
    """`get_model_cache_path` (`cache_paths.py:27-40`) without creating directories."""
    assert model_version, "model_version cannot be empty"
    assert file_name, "file_name cannot be empty"
    root = Path(cache_dir).expanduser().resolve() if cache_dir is not None else user_cache_dir()
    return root / model_version / file_name


def compute_sha256(file_path: str | Path, chunk_size: int = 1024 * 1024) -> str:
    """Streamed SHA256 of a file (`checksums.py:11-25`)."""
    path = Path(file_path)
    assert path.exists(), f"file does not exist: {path}"
    assert path.is_file(), f"path is not a file: {path}"
    h = hashlib.sha256()
    with path.open("rb") as stream:
        for chunk in iter(lambda: stream.read(chunk_size), b""):
            h.update(chunk)
    return h.hexdigest()


def assert_sha256(file_path: str | Path, expected_sha256: str) -> None:
    """ValueError on mismatch, same message as `checksums.py:38-45`."""
    assert expected_sha256, "expected_sha256 cannot be empty"
    actual = compute_sha256(file_path)
    if actual.lower() != expected_sha256.strip().lower():
        raise ValueError(f"checksum mismatch for {file_path}: expected {expected_sha256}, got {actual}")


def load_manifest(manifest_fp: str | Path | None = None) -> dict:
    """The `models.json` manifest: a file in the reference's format, else the built-in copy of its single entry."""
    if manifest_fp is None:
        return DEFAULT_MANIFEST
    data = json.loads(Path(manifest_fp).read_text(encoding="utf-8"))
    assert isinstance(data.get("models"), dict) and data["models"], "manifest must define a non-empty 'models' mapping"
    return data


def find_model(model_version: str = "ResUNet_16x_DEM", *, cache_dir: str | Path | None = None, manifest_fp: str | Path | None = None,
               inputs_dir: str | Path | None = None, verify: bool = True) -> Path | None:
    """Path of the cached weights of `model_version`, or None when they are not on disk.

    Order: `<inputs_dir>/<version>/*.onnx` (test convention), then the cache path.  A cached file whose digest does not match
    the manifest raises ValueError (the reference re-downloads in that case; there is no network here).
    """
    models = load_manifest(manifest_fp)["models"]
    assert model_version in models, f"unknown model version '{model_version}'; known: {sorted(models)}"
    entry = models[model_version]
    if inputs_dir is not None:
        hits = sorted((Path(inputs_dir) / model_version).glob("*.onnx"))
        if hits:
            return hits[0]
    fp = model_cache_path(model_version, entry["file_name"], cache_dir)
    if not fp.exists():
        return None
    if verify and entry.get("sha256"):
        assert_sha256(fp, entry["sha256"])
    return fp
