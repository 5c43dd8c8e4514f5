"""The header-only C++ shims (include/epivo_shims.hpp) compile against stand-in matrix/point types
and, on a GPU box, drive the C ABI like the reference drivers would."""
import os
import subprocess

import pytest

from epivo_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_test.bin")


def _compile():
    lib = build.build()
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
    cmd = ["g++", "-O1", "-std=c++11", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           lib, "-Wl,-rpath," + os.path.dirname(lib), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return EXE


def test_shims_compile_and_link():
    _compile()


@pytest.mark.gpu
def test_shims_run_like_a_reference_driver():
    exe = _compile()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "shims ok" in p.stdout
