"""the C-ABI library loads and exports every symbol include/ojph_b200.h declares (no GPU needed)"""
import os
import re
from openjph_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_bound():
    hdr = open(os.path.join(ROOT, "include", "ojph_b200.h")).read()
    declared = set(re.findall(r"\b(ojb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_product_library_exports_all_symbols():
    assert os.path.exists(_lib.LIB_PATH), "build libojph_b200.so first (__graft_entry__.build())"
    L = _lib.bind(_lib.LIB_PATH)
    assert L.ojb_version().startswith(b"openjph_b200")


def test_no_silent_fallback_when_library_missing(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    import pytest
    with pytest.raises(RuntimeError):
        _lib.lib()
