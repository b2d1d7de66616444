"""Weights discovery: same cache path, digest and manifest rules as the reference."""

from __future__ import annotations

import json
import sys

import pytest

from floodsr_b200 import model_store as ms


def test_cache_path_and_checksum_match_reference(tmp_path):
    fp = tmp_path / "blob.bin"
    fp.write_bytes(b"floodsr" * 1000)
    want_path = want_sha = None
    try:
        sys.path.insert(0, "/root/reference")
        from floodsr.cache_paths import get_model_cache_path
        from floodsr.checksums import compute_sha256

        want_path = get_model_cache_path("ResUNet_16x_DEM", "model_infer.onnx", cache_dir=tmp_path / "c")
        want_sha = compute_sha256(fp)
        default_root = get_model_cache_path("v", "f", cache_dir=None).parents[1]
        assert ms.user_cache_dir() == default_root
    except ImportError:
        pass  # the reference tree is not present on the GPU box
    finally:
        if "/root/reference" in sys.path:
            sys.path.remove("/root/reference")
    got_path = ms.model_cache_path("ResUNet_16x_DEM", "model_infer.onnx", cache_dir=tmp_path / "c")
    assert got_path == (tmp_path / "c").resolve() / "ResUNet_16x_DEM" / "model_infer.onnx"
    if want_path is not None:
        assert got_path == want_path and ms.compute_sha256(fp) == want_sha
    with pytest.raises(ValueError, match="checksum mismatch for .*: expected 00, got "):
        ms.assert_sha256(fp, "00")


def test_find_model_order_and_verification(tmp_path):
    sha = None
    cache = tmp_path / "cache"
    blob = cache / "ResUNet_16x_DEM" / "model_infer.onnx"
    assert ms.find_model(cache_dir=cache) is None
    blob.parent.mkdir(parents=True)
    blob.write_bytes(b"not the release asset")
    with pytest.raises(ValueError, match="checksum mismatch"):
        ms.find_model(cache_dir=cache)
    sha = ms.compute_sha256(blob)
    manifest = tmp_path / "models.json"
    manifest.write_text(json.dumps({"models": {"ResUNet_16x_DEM": {"file_name": "model_infer.onnx", "url": "file://x", "sha256": sha}}}))
    assert ms.find_model(cache_dir=cache, manifest_fp=manifest) == blob
    inputs = tmp_path / "_inputs"
    (inputs / "ResUNet_16x_DEM").mkdir(parents=True)
    (inputs / "ResUNet_16x_DEM" / "a.onnx").write_bytes(b"x")
    assert ms.find_model(cache_dir=cache, manifest_fp=manifest, inputs_dir=inputs).name == "a.onnx"
    with pytest.raises(AssertionError, match="unknown model version"):
        ms.find_model("CostGrow", cache_dir=cache)
    assert ms.DEFAULT_MANIFEST["models"]["ResUNet_16x_DEM"]["sha256"].startswith("ea907c0f")
