"""Compiles tests/cpp/dropin_test.cpp (a caller of the reference's C++ API, include/fpng.h) against libfpng_b200.so and
runs it on the GPU: the drop-in surface works for a C++ application without any Python in the loop."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin_program(tmp_path):
    from fpng_b200 import _build
    lib = _build.build()
    exe = str(tmp_path / "dropin_test")
    subprocess.check_call(["g++", "-O2", "-std=c++11", os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + os.path.dirname(lib), "-lfpng_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout
