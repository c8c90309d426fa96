"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a without a GPU, loads, and exports every
symbol include/fpng_b200.h declares.  No compute calls here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fpng_b200.h")).read()
    return sorted(set(re.findall(r"FPNGB_API[^;(]*?\b(fpngb_\w+)\s*\(", text)))


def test_library_builds_and_exports_declared_symbols():
    from fpng_b200 import _build
    path = _build.build()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/fpng_b200.h but not exported"


def test_cxx_dropin_header_symbols_exported():
    """include/fpng.h declares the reference's C++ API (namespace fpng); the library must define those too."""
    from fpng_b200 import _build
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", "-C", _build.build()], text=True)
    for name in ("fpng::fpng_init()", "fpng::fpng_encode_image_to_memory(", "fpng::fpng_decode_memory(", "fpng::fpng_get_info(",
                 "fpng::fpng_crc32(", "fpng::fpng_adler32(", "fpng::fpng_cpu_supports_sse41()", "fpng::fpng_encode_image_to_file(",
                 "fpng::fpng_decode_file("):
        assert name in out, name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fpng_b200
    from fpng_b200._lib import FpngB200Error
    with pytest.raises(FpngB200Error):
        fpng_b200.fpng_init()
    ok = None
    with pytest.raises(FpngB200Error):
        ok = fpng_b200.fpng_encode_image_to_memory(bytes(12), 2, 2, 3, 0)
    assert ok is None


def test_max_encoded_size_matches_oracle(oracle):
    import fpng_b200
    for (w, h, c) in ((1, 1, 3), (1, 1, 4), (512, 512, 4), (1920, 1080, 3), (3840, 2160, 4), (8193, 7, 3), (21845, 3, 3), (21846, 3, 3)):
        assert fpng_b200.max_encoded_size(w, h, c) == oracle.max_encoded_size(w, h, c)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fpng_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("// host twin", ""), f"{f} references oracle/"


def test_cxx_callers_compile_and_link_without_a_gpu(tmp_path):
    """The C++ programs that sit on the drop-in surface -- tests/cpp/dropin_test.cpp (a caller of include/fpng.h) and the
    fpng_test-equivalent harness tools/fpng_b200_test.cpp -- compile and link against libfpng_b200.so here; without a device they must
    refuse to run the codec (no CPU fallback): the harness exits non-zero."""
    import subprocess
    from fpng_b200 import _build
    lib = _build.build()
    exe = str(tmp_path / "dropin_test")
    subprocess.check_call(["g++", "-O1", "-std=c++11", os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + os.path.dirname(lib), "-lfpng_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "tools", "Makefile")])
    tool = os.path.join(ROOT, "tools", "fpng_b200_test")
    assert os.path.exists(tool)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([tool, os.path.join(ROOT, "tests", "golden", "example.png")], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0, r.stdout[-300:]
