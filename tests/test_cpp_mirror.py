"""The C++ host mirror (include/spalinalg.hpp): the typed host side above the C ABI in the
language class of the reference (compiled code; no Rust toolchain in the image).
CPU part: the header compiles with g++ -std=c++17 and the test program links against the in-tree
library (every ABI symbol it uses resolves).  GPU part: the program — the reference's own hot-path
unit tests restated in C++ — runs and passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "spalinalg_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_mirror")


def _build():
    from spalinalg_b200 import build as spl_build
    spl_build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", EXE, "-L", LIBDIR,
           "-lspalinalg_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_passes_reference_tests():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
