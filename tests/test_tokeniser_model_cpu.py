"""CPU model of the data-parallel tokeniser the 16-pixel CUDA kernels implement (csrc/row_walk16.cuh classify16 +
lane_info16, csrc/encode16_kernels.cu run_before16 / put_event16), checked against the serial restatement of the
reference's tokeniser (fpng.cpp:1182-1243 / 1468-1558; oracle/fpng_oracle.c tokenise_row) -- which is itself pinned to the
oracle here through the per-row token bit counts `oracle_encode_ex` reports.

This is host logic (no GPU): it documents WHY 32 lanes of 16 pixels can emit their tokens independently.  A run of pixels
equal to their left neighbour is cut into matches of M pixels from the run start; a token belongs to the pixel where it
ends (full match), or to the literal / row end that follows it (remainder).  Per lane the kernels need only
  eqm  -- 16-bit mask "pixel k equals its left neighbour",
  r_in -- the unfinished run entering the lane (from the nearest lower lane holding a literal: one ballot + one shuffle),
and derive from them (a) the literal pixels, (b) the "events": a pending run flushed before a literal, or the run
reaching M at a match pixel (at most one per lane since M > 16)."""
import numpy as np
import pytest

import imagegen

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_XBITS = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def len_symbol(L):
    """RFC 1951 length code (symbol, extra bits) of a match of L bytes (fpng.cpp:498-512)."""
    if L == 258:
        return 285, 0
    s = max(i for i in range(28) if LEN_BASE[i] <= L)
    return 257 + s, LEN_XBITS[s]


def filtered_pixels(img, y):
    """PNG filter 0 on row 0, filter 2 (Up) below (fpng.cpp:1592-1660); returns (filter byte, [w] packed pixels)."""
    h, w, c = img.shape
    row = img[y].astype(np.int16)
    d = row if y == 0 else (row - img[y - 1].astype(np.int16)) & 255
    d = d.astype(np.uint32)
    px = d[:, 0] | (d[:, 1] << 8) | (d[:, 2] << 16)
    if c == 4:
        px = px | (d[:, 3] << 24)
    return (0 if y == 0 else 2), [int(v) for v in px]


def serial_tokens(filt, px, M):
    out = [("F", filt)]
    prev, run = None, 0
    for x, p in enumerate(px):
        if x and p == prev:
            run += 1
            if run == M:
                out.append(("M", M)); run = 0
        else:
            if run:
                out.append(("M", run)); run = 0
            out.append(("L", p)); prev = p
    if run:
        out.append(("M", run))
    return out


def run_before(eqm, r_in, kk, M):
    """Pending run length before pixel kk of a lane (encode16_kernels.cu run_before16)."""
    t = ~eqm & ((1 << kk) - 1)
    if t == 0:
        r = kk + r_in
        return r - M if r >= M else r
    return kk - 1 - (t.bit_length() - 1)


def parallel_tokens(filt, px, M):
    """Tokens of one scanline as 32 lanes x 16 pixels per warp step produce them, lanes concatenated in order."""
    w = len(px)
    out = [("F", filt)]
    carry_prev, carry_run = 0, 0
    for s0 in range(0, w, 512):
        lanes = []
        for lane in range(32):
            p0 = s0 + lane * 16
            nvp = min(16, w - p0) if p0 < w else 0
            eqm = 0
            for k in range(nvp):
                x = p0 + k
                left = px[x - 1] if x > 0 else None
                if x > 0 and px[x] == left:
                    eqm |= 1 << k
            valid = (1 << nvp) - 1
            litm = valid & ~eqm
            trail = (nvp - 1 - (litm.bit_length() - 1)) if litm else nvp
            lanes.append(dict(p0=p0, nvp=nvp, eqm=eqm, litm=litm, trail=trail))
        # run phase entering each lane: the nearest lower lane with a literal decides (ballot + shuffle in classify16)
        for lane, t in enumerate(lanes):
            src = max((l for l in range(lane) if lanes[l]["litm"]), default=None)
            run = (lanes[src]["trail"] + 16 * (lane - src - 1)) if src is not None else (carry_run + 16 * lane)
            t["run"] = run % M
        last = lanes[31]
        carry_run = last["trail"] if last["litm"] else (last["run"] + last["nvp"]) % M
        # per-lane emission from (eqm, r_in) alone: literals + events
        for t in lanes:
            eqm, litm, nvp, r_in = t["eqm"], t["litm"], t["nvp"], t["run"]
            lead = ((~eqm) & -(~eqm)).bit_length() - 1            # leading match pixels (ffs(~eqm) - 1)
            kM = M - 1 - r_in
            evm = (litm & ((eqm << 1) | (1 if r_in else 0))) | ((1 << kM) if kM < lead else 0)
            for k in range(nvp):
                if evm >> k & 1:
                    n = M if eqm >> k & 1 else run_before(eqm, r_in, k, M)
                    if n:
                        out.append(("M", n))
                if litm >> k & 1:
                    out.append(("L", px[t["p0"] + k]))
            if nvp and t["p0"] + nvp == w:                         # the lane holding the scanline's last pixel
                n = run_before(eqm, r_in, nvp, M)
                if n:
                    out.append(("M", n))
    return out


def token_bits(tokens, sizes, chans):
    bits = 0
    for kind, v in tokens:
        if kind == "F":
            bits += int(sizes[v])
        elif kind == "L":
            bits += sum(int(sizes[(v >> (8 * c)) & 255]) for c in range(chans))
        else:
            sym, xb = len_symbol(v * chans)
            bits += int(sizes[sym]) + xb + 1                       # + the 1-bit distance code (fpng.cpp:1135)
    return bits


@pytest.mark.parametrize("chans", [3, 4])
@pytest.mark.parametrize("kind", ["g1", "g0", "runs", "mut", "zero"])
def test_lane_parallel_tokeniser_equals_serial_and_oracle(oracle, kind, chans):
    M = 85 if chans == 3 else 63
    sizes, _, _ = oracle.static_table(chans)
    for i, (w, h) in enumerate([(1, 3), (17, 2), (85, 3), (512, 4), (513, 3), (700, 5), (1100, 3), (2049, 2)]):
        img = imagegen.make(kind, w, h, chans, 7 + i)
        _, row_bits, stored = oracle.encode(img, w, h, chans, 0, want_rows=True)
        for y in range(h):
            filt, px = filtered_pixels(img, y)
            ser = serial_tokens(filt, px, M)
            par = parallel_tokens(filt, px, M)
            assert par == ser, (kind, chans, w, h, y)
            if not stored:   # the oracle reports each row's token bits of the compressed attempt
                assert token_bits(ser, sizes, chans) == int(row_bits[y]), (kind, chans, w, h, y)


def test_run_boundaries_at_max_match():
    """Runs that cross lanes, steps and the maximum match length M (255/3 and 252/4 pixels)."""
    for chans, M in ((3, 85), (4, 63)):
        for run_len in (M - 1, M, M + 1, 2 * M, 2 * M + 5, 511, 512, 513, 1030):
            for lead_in in (0, 1, 15, 16, 17, 500):
                px = list(range(1, lead_in + 1)) + [777] * run_len + [5, 5, 6]
                assert parallel_tokens(2, px, M) == serial_tokens(2, px, M), (chans, run_len, lead_in)
