"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/epivo_b200.h
declares (no compute calls: this runs without a GPU)."""
import os
import re

from epivo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "epivo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(epivo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/epivo_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"


def test_version_string():
    assert b"sm_100a" in _lib.load().epivo_version()


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_lib.LmRes) == 24
    assert C.sizeof(_lib.PairResult) == (9 + 9 + 3 + 16 + 16 + 3) * 8 + 8 * 4
