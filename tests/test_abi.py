"""CPU-only: the C-ABI library builds, loads without a GPU, and exports exactly what include/segvlad.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from revisit_anything_b200 import build
    return build.build()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "segvlad.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(segvlad_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols(built):
    h = ctypes.CDLL(built)
    names = _header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/segvlad.h but not exported"


def test_python_binding_matches_header(built):
    from revisit_anything_b200 import _lib
    assert sorted(_lib.EXPORTED) == _header_functions()
    lib = _lib.lib()
    assert lib.segvlad_version() >= 100


def test_workspace_queries_are_host_only(built):
    from revisit_anything_b200 import _lib
    lib = _lib.lib()
    assert lib.segvlad_aggregate_workspace_bytes(4, 1530, 1536, 32, 500) > 4 * 1530 * 1536 * 4
    assert lib.segvlad_bank_bytes(1000, 1536) >= 1000 * 1536 * 4
    assert lib.segvlad_bank_bytes(1000, 100) >= 1000 * 128 * 4      # padded to 64 columns
    assert lib.segvlad_knn_workspace_bytes(10000, 100000, 1536, 200) > 10000 * 4096 * 8
    assert lib.segvlad_vote_workspace_bytes(1000, 50, 10, 100) >= 256


def test_missing_library_fails_loudly(monkeypatch, built):
    from revisit_anything_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsegvlad.so")
    with pytest.raises(_lib.SegVladError):
        _lib.lib()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "revisit-anything_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("checker", ""), f"{fn} references the oracle"
