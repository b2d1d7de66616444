"""Structural guards: the product never touches the oracle or the reference tree, and fails loudly without its library."""
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
PRODUCT = sorted((REPO / "floodsr_b200").rglob("*.py")) + sorted((REPO / "floodsr_b200" / "csrc").glob("*"))


def test_product_path_never_imports_the_oracle_or_reads_the_reference_tree():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    offenders = [str(p.relative_to(REPO)) for p in PRODUCT if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h"} and pat.search(p.read_text())]
    assert offenders == []


def test_only_the_allowed_entry_points_use_the_oracle():
    users = []
    for p in [REPO / "bench.py", REPO / "__graft_entry__.py"] + sorted((REPO / "scripts").rglob("*.py")):
        if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), re.M):
            users.append(p.name)
    # bench.py: cpu_baseline / --impl reference legs; __graft_entry__.py: smoke() as the checker
    assert set(users) <= {"bench.py", "__graft_entry__.py", "make_golden.py"}, users


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    from floodsr_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libfloodsr_b200_missing.so")
    with pytest.raises(_lib.EngineLibraryError, match="no CPU fallback"):
        _lib.load_library()
