"""The fpng_test-equivalent command-line harness (tools/fpng_b200_test.cpp; reference: src/fpng_test.cpp:975-1639) on the GPU:
default run, -s, -u, -a, -c (CSV), -f, -e and -E with a few trials, and the general-PNG fallback hook on a file that is
not fpng-written (FPNG_DECODE_NOT_FPNG -> registered decoder)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import imagegen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tools", "fpng_b200_test")
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example.png")


@pytest.fixture(scope="module")
def exe():
    from fpng_b200 import _build
    _build.build()
    subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "tools", "Makefile")])
    return EXE


def run(exe, *args, cwd=None):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, cwd=cwd)


def write_general_png(path, img):
    """A PNG no fpng decoder accepts: no fdEC chunk, Paeth/Sub/Average filters, ordinary zlib stream."""
    h, w, c = img.shape
    raw = bytearray()
    prev = np.zeros((w, c), np.int16)
    for y in range(h):
        cur = img[y].astype(np.int16)
        ft = 1 + y % 3                                      # Sub, Up, Average
        left = np.vstack([np.zeros((1, c), np.int16), cur[:-1]])
        pred = left if ft == 1 else prev if ft == 2 else (left + prev) // 2
        raw.append(ft)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(bytes(raw), 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("opts", [[], ["-s"], ["-u"], ["-a"], ["-a", "-s"]])
def test_cli_on_example_png(exe, ref, tmp_path, opts):
    out_png = str(tmp_path / "fpng.png")
    r = run(exe, *opts, "-o", out_png, EXAMPLE)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Verified" in r.stdout
    data = open(out_png, "rb").read()
    # the written file is what the reference writes for the same pixels/flags
    err, px, w, h = ref.lodepng_decode(open(EXAMPLE, "rb").read(), 3)
    flags = (1 if "-s" in opts else 0) | (2 if "-u" in opts else 0)
    if "-a" in opts:
        rgba = np.concatenate([px.reshape(-1, 3), px.reshape(-1, 3)[:, 1:2]], axis=1).reshape(-1)
        assert data == ref.encode(rgba, w, h, 4, flags)
    else:
        assert data == ref.encode(px, w, h, 3, flags)


def test_cli_csv_and_decode_only(exe, tmp_path):
    r = run(exe, "-c", "-o", str(tmp_path / "o.png"), EXAMPLE)
    assert r.returncode == 0, r.stderr
    f = [t.strip() for t in r.stdout.strip().splitlines()[-1].split(",")]
    assert f[0] == EXAMPLE and f[1:4] == ["687", "1012", "3"] and len(f) == 9 and float(f[7]) > 0 and float(f[8]) > 0
    assert run(exe, "-f", EXAMPLE).returncode == 0
    bad = tmp_path / "bad.png"
    d = bytearray(open(EXAMPLE, "rb").read()); d[18] ^= 0x55          # IHDR width: header CRC-32 mismatch (the IDAT CRC is not checked by fpng decoders)
    bad.write_bytes(bytes(d))
    assert run(exe, "-f", str(bad)).returncode != 0
    bad.write_bytes(open(EXAMPLE, "rb").read()[:-40])                    # truncated stream
    assert run(exe, "-f", str(bad)).returncode != 0


def test_cli_general_png_goes_through_the_fallback_hook(exe, gpu, tmp_path):
    img = imagegen.make("g1", 200, 37, 4, 5)
    src = str(tmp_path / "general.png")
    write_general_png(src, img)
    assert gpu.fpng_get_info(open(src, "rb").read())[0] == gpu.FPNG_DECODE_NOT_FPNG
    assert run(exe, "-f", src).returncode != 0                       # raw fpng status: not an fpng file
    out_png = str(tmp_path / "o.png")
    r = run(exe, "-o", out_png, src)                                    # loader = fpng_decode_memory + registered fallback decoder
    assert r.returncode == 0, r.stdout + r.stderr
    st, px, w, h, c = gpu.fpng_decode_memory(open(out_png, "rb").read(), 4)
    assert st == 0 and (w, h, c) == (200, 37, 4) and np.array_equal(px, img.reshape(-1))


def test_cli_fuzz_modes(exe):
    r = run(exe, "-e", "-n", "6", EXAMPLE)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    sizes = [int(l.split(":")[1]) for l in r.stdout.splitlines() if l.startswith("fpng size")]
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "fpng_test_fuzz.json")))
    assert sizes == g["e_sizes"][:6]                                   # same numbers the reference harness prints
    r = run(exe, "-E", "-n", "3")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    got = [l for l in r.stdout.splitlines() if l.startswith(("Testing", "fpng size"))]
    exp = []
    for w, h, c, s in g["E_trials"][:3]:
        exp += [f"Testing {w}x{h} {c}", f"fpng size: {s}"]
    assert got == exp
