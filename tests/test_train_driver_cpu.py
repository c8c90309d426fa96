"""tools/train_tables.py (the `fpng_test -t` counterpart) against the REAL harness: the unmodified reference's fpng_test is compiled
with its own FPNG_TRAIN_HUFFMAN_TABLES=1 switch (into oracle/_ref/, test-only) and run on a listing of PNG files; the driver -- with
the reference trainer (oracle/_ref/libfpng_ref_train.so) standing in for the GPU backend, since this test has no device -- must print
the same text: per-file lines (dimensions, alpha detection src/fpng_test.cpp:808-821), totals, and both C tables character for
character (src/fpng_test.cpp:893-961).  The GPU backend's two calls (histogram accumulation, prefix builder) are pinned to the same
reference trainer by tests/test_training_gpu.py."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

import imagegen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
HARNESS = os.path.join(ROOT, "oracle", "_ref", "fpng_test_train")


@pytest.fixture(scope="module")
def harness():
    if not os.path.exists(HARNESS):
        if not os.path.exists(os.path.join(REF_SRC, "fpng_test.cpp")):
            pytest.skip("reference tree not present and no prebuilt harness")
        os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++11", "-fno-strict-aliasing", "-msse4.1", "-mpclmul", "-DFPNG_NO_SSE=0",
                               "-DFPNG_TRAIN_HUFFMAN_TABLES=1", "-w", "-I" + REF_SRC] +
                              [os.path.join(REF_SRC, f) for f in ("fpng.cpp", "fpng_test.cpp", "lodepng.cpp", "pvpngreader.cpp")] +
                              ["-o", HARNESS, "-lm", "-lpthread"])
    return HARNESS


class ReferenceTrainerBackend:
    """stands in for tools/train_tables.py GpuBackend (same three calls) on a box without a device"""

    def __init__(self):
        from oracle.pyoracle import RefTrainer, Ref
        self.t, self.ref = RefTrainer(), Ref()

    def accumulate(self, img, counts):
        h, w, c = img.shape
        counts += self.t.counts_from_encodes([img], w, h, c)

    def create_prefix(self, counts, chans):
        return self.t.create_prefix(counts, chans)

    def sanity(self, img):
        h, w, c = img.shape
        st, px, *_ = self.ref.decode(self.ref.encode(img, w, h, c, 1), c)
        return st == 0 and np.array_equal(px, img.reshape(-1))


def _write_pngs(tmp, oracle):
    names = []
    cases = [("g1", 96, 40, 3), ("g1", 64, 64, 4), ("g0", 200, 30, 4), ("runs", 150, 20, 3), ("g2", 31, 17, 3), ("mut", 120, 33, 4),
             ("g1", 50, 50, 4)]
    for i, (kind, w, h, c) in enumerate(cases):
        img = np.asarray(imagegen.make(kind, w, h, c, i)).reshape(h, w, c).copy()
        if i == 6:
            img[:, :, 3] = 255                                  # an RGBA file whose alpha is all opaque trains the 24bpp table (Q8)
        p = os.path.join(tmp, f"img{i}.png")
        with open(p, "wb") as f:
            f.write(oracle.encode(img, w, h, c, 0))
        names.append(p)
    bad = os.path.join(tmp, "broken.png")
    with open(bad, "wb") as f:
        f.write(b"not a png at all")
    names.insert(3, bad)                                        # skipped with a warning, counted as failed
    listing = os.path.join(tmp, "list.txt")
    with open(listing, "w") as f:
        f.write("\n".join(names) + "\n\n")
    return listing


def test_training_driver_prints_what_the_real_harness_prints(tmp_path, oracle, harness):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_tables
    listing = _write_pngs(str(tmp_path), oracle)
    real = subprocess.run([harness, "-t", "@" + listing], capture_output=True, text=True)
    assert real.returncode == 0, real.stderr
    # the harness prints a banner / SSE line first; compare from the listing line on
    real_lines = real.stdout.splitlines()
    start = next(i for i, l in enumerate(real_lines) if l.startswith("Successfully read"))
    buf = io.StringIO()
    files = train_tables.load_listing("@" + listing)
    assert len(files) == 8
    tables = train_tables.train(files, ReferenceTrainerBackend(), out=buf)
    mine = [real_lines[start]] + buf.getvalue().splitlines()
    assert mine == real_lines[start:], "\n".join(f"{a!r}\n{b!r}" for a, b in zip(mine, real_lines[start:]) if a != b)
    assert "Total alpha files: 3" in mine and "Total opaque files: 4" in mine and "Total failed loading: 1" in mine
    assert tables[3] is not None and tables[4] is not None and len(tables[3][0]) > 40
