"""CPU: the C-ABI library loads and exports every symbol include/kodb.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    lib = os.path.join(ROOT, "object_detection_cib_b200", "libkodb.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "object_detection_cib_b200", "csrc"), "-j8"])
    return lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kodb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kodb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built):
    from object_detection_cib_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/kodb.h"
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes binding covers exactly the header
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kodb_version() >= 100
    assert lib.kodb_error_string(-2).decode().startswith("capacity")


def test_missing_library_fails_loudly(monkeypatch, built):
    from object_detection_cib_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libkodb.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.lib()
