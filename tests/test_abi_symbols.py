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


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (what a cgo / FFI binding would include)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    src = tmp_path / "use_header.c"
    src.write_text('#include "epivo_b200.h"\n'
                   "int main(void) { epivo_pipeline_params p; epivo_pair_result r; epivo_lm_res l; "
                   "(void)p; (void)r; (void)l; return sizeof(epivo_pair_result) == %d ? 0 : 1; }\n"
                   % ((9 + 9 + 3 + 16 + 16 + 3) * 8 + 8 * 4))
    exe = tmp_path / "use_header"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0          # same struct size as the ctypes mirror


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under epivo_b200/ (nor the tuning tools) may import or execute it --
    only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm do."""
    import ast
    offenders = []
    for top in ("epivo_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                path = os.path.join(dirpath, f)
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                        offenders.append(os.path.relpath(path, ROOT))
    assert not offenders, offenders
    for dirpath, _, files in os.walk(os.path.join(ROOT, "epivo_b200", "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                assert "oracle/" not in open(os.path.join(dirpath, f)).read(), f
