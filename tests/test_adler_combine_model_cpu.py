"""CPU model of the encoder's Adler-32 (reference: one fpng_adler32 call over the whole filtered stream, src/fpng.cpp:1150-1158,
403-487): the scan kernel leaves per-scanline partial sums, built from per-lane dot products (csrc/row_walk16.cuh adler_chunk16 /
adler_lane16: sum of bytes and position-weighted sum, one `dp4a` per 4 bytes), and csrc/encode_kernels.cu adler_finalize_kernel
combines the scanlines in closed form -- no serial pass over the 6-33 MB stream.

For a block X of n bytes appended to the state (a, b):  a' = a + S1,  b' = b + n*a + S2  (mod 65521) with
S1 = sum x_i, S2 = sum (n - i) x_i.  All scanline blocks have the same length n = bpl + 1 (filter byte first), so over h scanlines
    a = 1 + sum_y S1_y,     b = sum_y S2_y + n * (h + sum_y (h - 1 - y) * S1_y),
which is what the kernel evaluates with three block reductions."""
import zlib

import numpy as np
import pytest

import imagegen

MOD = 65521


def filtered_stream(img):
    h, w, c = img.shape
    rows = np.zeros((h, w * c + 1), np.uint8)
    flat = img.reshape(h, w * c)
    rows[0, 1:] = flat[0]
    rows[1:, 0] = 2
    rows[1:, 1:] = flat[1:] - flat[:-1]                                   # uint8 arithmetic wraps: PNG filter 2 (Up), fpng.cpp:1605-1652
    return rows


def lane_partials(data_bytes, chans):
    """what a warp accumulates for one scanline's data bytes: lanes own 16 pixels (16 * chans bytes) per 512-pixel step; per lane
    t1 = sum of its bytes, t2 = sum of (position inside the lane) * byte; the scanline keeps sumA = sum t1 and
    sumB = sum (lane_base * t1 + t2) = sum over data bytes of (byte index) * byte"""
    lane_bytes = 16 * chans
    sumA = sumB = 0
    for base in range(0, len(data_bytes), lane_bytes):
        chunk = data_bytes[base:base + lane_bytes].astype(np.int64)
        t1 = int(chunk.sum())
        t2 = int((np.arange(len(chunk)) * chunk).sum())                     # the dp4a weights 0, 1, 2, ... (kW in adler_chunk16)
        sumA += t1
        sumB += base * t1 + t2
    return sumA, sumB


@pytest.mark.parametrize("kind,w,h,chans", [("g1", 97, 13, 3), ("g1", 64, 9, 4), ("g2", 700, 5, 3), ("g0", 1030, 3, 4), ("zero", 5, 4, 3)])
def test_scanline_partials_combine_to_adler32(oracle, kind, w, h, chans):
    img = np.asarray(imagegen.make(kind, w, h, chans, 2)).reshape(h, w, chans)
    rows = filtered_stream(img)
    n = rows.shape[1]
    S1, S2 = [], []
    for y in range(h):
        sumA, sumB = lane_partials(rows[y, 1:], chans)
        filt = int(rows[y, 0])
        # data byte j sits at index 1 + j of the scanline block: weight n - (1 + j); the filter byte has weight n
        s1 = filt + sumA
        s2 = n * filt + (n - 1) * sumA - sumB
        assert s1 == int(rows[y].astype(np.int64).sum())
        assert s2 == int(((n - np.arange(n)) * rows[y].astype(np.int64)).sum())
        S1.append(s1 % MOD); S2.append(s2 % MOD)
    a = (1 + sum(S1)) % MOD
    before = (h + sum((h - 1 - y) * S1[y] for y in range(h))) % MOD
    b = (sum(S2) + n * before) % MOD
    adler = (b << 16) | a
    assert adler == zlib.adler32(rows.tobytes())
    # ... and it is the value in the file (big-endian, the last 4 bytes of the zlib stream); stored files checksum the filter-0 stream
    for flags in (0, 2):
        png = oracle.encode(img, w, h, chans, flags)
        zsize = int.from_bytes(png[50:54], "big")
        if (png[60] & 6) == 0:                 # stored blocks: forced, or the incompressible fallback (fpng.cpp:1728-1758)
            stored = rows.copy(); stored[:, 0] = 0; stored[1:, 1:] = img.reshape(h, w * chans)[1:]     # stored files: filter 0 on every scanline
            assert int.from_bytes(png[58 + zsize - 4:58 + zsize], "big") == zlib.adler32(stored.tobytes())
        else:
            assert int.from_bytes(png[58 + zsize - 4:58 + zsize], "big") == adler
