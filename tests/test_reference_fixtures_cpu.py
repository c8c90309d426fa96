"""CPU side of the reference's own fixtures: example.png (the only file the reference ships; fpng_test's default input,
src/fpng_test.cpp:1118) and the inputs of its two encoder fuzzers (-e, -E) regenerated with the reference's seeds.
Pins (a) the oracle's decoder on a file written by an OLDER fpng than the one restated, (b) tests/cpp/fuzzgen.cpp against
sizes captured from the real harness (tests/golden/fpng_test_fuzz.json)."""
import json
import os

import numpy as np
import pytest

import fuzzgen
from common import HERE, sha

EXAMPLE = os.path.join(HERE, "golden", "example.png")

def example_bytes():
    with open(EXAMPLE, "rb") as f:
        return f.read()


def fuzz_golden():
    with open(os.path.join(HERE, "golden", "fpng_test_fuzz.json")) as f:
        return json.load(f)


def test_example_png_oracle_decode_matches_reference_and_lodepng(oracle, ref):
    data = example_bytes()
    assert len(data) == 1479224
    assert oracle.get_info(data) == (0, 687, 1012, 3) == ref.get_info(data)
    err, lode3, w, h = ref.lodepng_decode(data, 3)
    assert err == 0 and (w, h) == (687, 1012)
    for desired in (3, 4):
        st_o, px_o, *dims_o = oracle.decode(data, desired)
        st_r, px_r, *dims_r = ref.decode(data, desired)
        assert st_o == 0 == st_r and dims_o == dims_r == [687, 1012, 3]
        assert np.array_equal(px_o, px_r)
        if desired == 3:
            assert np.array_equal(px_o, lode3)
        else:
            q = px_o.reshape(-1, 4)
            assert np.array_equal(q[:, :3].reshape(-1), lode3) and (q[:, 3] == 255).all()
    comp, stb, *_ = ref.stb_decode(data, 3)
    assert comp and np.array_equal(stb, lode3)


@pytest.mark.parametrize("flags", [0, 1])
def test_example_png_reencode_oracle_equals_reference(oracle, ref, flags):
    """fpng_test's main flow on its default input: decode with lodepng, encode 1-pass / -s (src/fpng_test.cpp:1200-1209).
    687 x 3 = 2061-byte scanlines: no alignment of any kind."""
    err, px, w, h = ref.lodepng_decode(example_bytes(), 3)
    assert err == 0
    png = oracle.encode(px, w, h, 3, flags)
    assert png == ref.encode(px, w, h, 3, flags)
    assert len(png) == (1547277 if flags == 0 else 1479225)          # the v1.0.6 `-s` re-encode is one byte longer than the shipped file


def test_fuzzgen_e_matches_real_harness(oracle, ref):
    g = fuzz_golden()
    err, px, w, h = ref.lodepng_decode(example_bytes(), 3)
    trials = [0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987]
    trials += [i for i, f in enumerate(g["e_family"]) if f in (0, 1, 2, 4)][:12]           # the rare families too
    for t in sorted(set(trials)):
        buf, fam = fuzzgen.mutate(t, px, 3)
        assert fam == g["e_family"][t], t
        png = oracle.encode(buf, w, h, 3, 0)
        assert len(png) == g["e_sizes"][t], (t, fam)
        if t % 3 == 0:
            assert png == ref.encode(buf, w, h, 3, 0)


def test_fuzzgen_E_matches_real_harness(ref):
    g = fuzz_golden()
    s = fuzzgen.DimSession()
    try:
        for t, (w, h, c, size) in enumerate(g["E_trials"][:10]):
            small = w * h <= 3_000_000
            ww, hh, cc, buf = s.next(want_pixels=small)
            assert (ww, hh, cc) == (w, h, c), t
            if small:
                assert len(ref.encode(buf, w, h, c, 0)) == size
    finally:
        s.close()


def test_five_decoder_round_trip_of_oracle_and_reference_files(oracle, ref, verifiers):
    """The reference's integration check (src/fpng_test.cpp:1237-1445, 1571-1606: fpng, lodepng, stb_image, wuffs, pvpng all return
    the source pixels) on files written by the oracle and by the unmodified reference -- the GPU encoder's files are byte-identical
    to these (tests/test_encode_gpu.py), and its own five-decoder test runs under `-m gpu`.  wuffs verifies both checksums here."""
    import imagegen
    for i, (kind, w, h, c) in enumerate([("g1", 301, 57, 3), ("g1", 300, 57, 4), ("runs", 1024, 9, 4), ("g2", 77, 31, 3), ("g0", 640, 48, 4),
                                          ("mut", 333, 21, 3), ("zero", 1, 1, 4), ("g1", 1, 9, 3), ("g1", 8194, 2, 3)]):
        img = np.asarray(imagegen.make(kind, w, h, c, i)).reshape(h, w, c)
        rgba = img if c == 4 else np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], axis=2)
        for flags in (0, 1, 2):
            png = oracle.encode(img, w, h, c, flags)
            assert png == ref.encode(img, w, h, c, flags)
            st, px, ww, hh, cc = ref.decode(png, c)
            assert st == 0 and np.array_equal(px, img.reshape(-1))
            err, px, ww, hh = ref.lodepng_decode(png, c)
            assert err == 0 and np.array_equal(px, img.reshape(-1))
            comp, px, ww, hh = ref.stb_decode(png, c)
            assert comp == c and np.array_equal(px, img.reshape(-1))
            rc, px, ww, hh = verifiers.wuffs_decode_rgba(png, w, h)
            assert rc == 0 and (ww, hh) == (w, h) and np.array_equal(px, rgba.reshape(-1)), (kind, w, h, c, flags)
            rc, px, ww, hh, cc = verifiers.pvpng_decode(png, c, w, h)
            assert rc == 0 and (ww, hh, cc) == (w, h, c) and np.array_equal(px, img.reshape(-1)), (kind, w, h, c, flags)
            # a flipped checksum bit is caught by wuffs (IDAT CRC-32: 16 bytes from the end; Adler-32: the 4 bytes before it)
            for pos in (len(png) - 13, len(png) - 17):
                bad = bytearray(png); bad[pos] ^= 0x10
                assert verifiers.wuffs_decode_rgba(bytes(bad), w, h)[0] == 1
