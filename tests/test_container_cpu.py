"""Host-side container walk of the product library (fpngb_get_info / fpngb_get_info_ex; reference: fpng_get_info,
src/fpng.cpp:2930-3083) checked on the CPU against the unmodified reference and the oracle.  The walk is plain host code
in libfpng_b200.so (signature, IHDR, fdEC, IDAT bookkeeping, chunk CRCs): no kernel is launched, so this runs without a GPU."""
import numpy as np
import pytest

import imagegen


def files(oracle):
    out = []
    for i, (kind, w, h, c, flags) in enumerate([("g1", 40, 6, 3, 0), ("g0", 200, 20, 4, 0), ("g1", 64, 16, 3, 1), ("g2", 31, 9, 4, 0),
                                                 ("runs", 300, 7, 3, 2), ("zero", 1, 1, 4, 0), ("g1", 1025, 3, 4, 1)]):
        img = imagegen.make(kind, w, h, c, i)
        out.append((oracle.encode(img, w, h, c, flags), w, h, c))
    return out


def test_get_info_on_valid_files(oracle):
    import fpng_b200
    for png, w, h, c in files(oracle):
        assert fpng_b200.fpng_get_info(png) == (0, w, h, c)
        st, ww, hh, cc, ofs, ln = fpng_b200.get_info_ex(png)
        assert (st, ww, hh, cc) == (0, w, h, c)
        # idat_ofs is the IDAT chunk's start (its 4-byte big-endian length, then "IDAT", then idat_len payload bytes)
        assert png[ofs + 4:ofs + 8] == b"IDAT" and int.from_bytes(png[ofs:ofs + 4], "big") == ln
        assert oracle.get_info(png)[0] == 0


def test_get_info_on_corrupted_containers_matches_reference(oracle, ref):
    import fpng_b200
    rs = np.random.RandomState(11)
    n = 0
    for png, w, h, c in files(oracle):
        cases = [png[:k] for k in (0, 7, 8, 20, 33, 40, 57, 58, len(png) - 13, len(png) - 12, len(png) - 1)]
        for _ in range(150):                      # bit flips, mostly in the chunk structure (first 100 and last 24 bytes)
            bad = bytearray(png)
            zone = int(rs.randint(0, 3))
            pos = int(rs.randint(0, min(100, len(bad)))) if zone == 0 else (len(bad) - 1 - int(rs.randint(0, min(24, len(bad)))) if zone == 1
                                                                             else int(rs.randint(0, len(bad))))
            bad[pos] ^= 1 << int(rs.randint(0, 8))
            cases.append(bytes(bad))
        for _ in range(20):                       # spliced garbage / duplicated chunks
            cut = int(rs.randint(8, len(png)))
            cases.append(png[:cut] + bytes(rs.randint(0, 256, int(rs.randint(1, 40)), dtype=np.uint8)) + png[cut:])
        for bad in cases:
            if len(bad) == 0:
                continue
            exp = ref.get_info(bad)
            got = fpng_b200.fpng_get_info(bad)
            assert got[0] == exp[0], (len(bad), got, exp)
            if exp[0] == 0:
                assert got[1:] == tuple(exp[1:4]), (got, exp)
            n += 1
    assert n > 1000


def test_product_static_code_books_match_oracle(oracle):
    """The 1-pass code books the product derives at init by parsing the two pre-serialised block headers
    (csrc/static_tables.h; reference: fpng.cpp:184-371 g_dyn_huff_3/4 + their size/code tables) -- host code, no GPU."""
    import ctypes as C
    from fpng_b200 import _lib
    L = _lib.lib()
    for chans in (3, 4):
        sizes = np.zeros(288, np.uint8); codes = np.zeros(288, np.uint16); hb = C.c_uint32(0)
        rc = L.fpngb_debug_static_table(chans, sizes.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p), C.byref(hb))
        assert rc == 0
        esz, ecd, ehb = oracle.static_table(chans)
        assert np.array_equal(sizes, esz) and hb.value == ehb
        assert np.array_equal(codes[:257], ecd[:257])          # literal codes + end of block (the hook reports no length codes)
        # ... and, independently of the oracle, the tables the reference itself stores (tests/golden/static_tables.json, read from
        # the reference source text by tests/golden/make_static_tables.py: fpng.cpp:532-562 and the RGBA pair below it)
        import json, os
        g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "static_tables.json")))[str(chans)]
        assert list(sizes) == g["sizes"]
        assert all(codes[i] == g["codes"][i] for i in range(257) if g["sizes"][i])
        assert hb.value == 8 * len(g["header_bytes"]) + g["bit_buf_size"]
        assert np.array_equal(esz, g["sizes"]) and all(ecd[i] == g["codes"][i] for i in range(288) if g["sizes"][i])


def _chunk(tag: bytes, body: bytes, good_crc: bool = True) -> bytes:
    import zlib
    crc = zlib.crc32(tag + body) & 0xFFFFFFFF
    if not good_crc:
        crc ^= 0x5A5A5A5A
    return len(body).to_bytes(4, "big") + tag + body + crc.to_bytes(4, "big")


def _split(png: bytes):
    """signature, [(tag, body, raw chunk bytes)] of a well-formed PNG"""
    out, at = [], 8
    while at < len(png):
        ln = int.from_bytes(png[at:at + 4], "big")
        out.append((png[at + 4:at + 8], png[at + 8:at + 8 + ln], png[at:at + 12 + ln]))
        at += 12 + ln
    return png[:8], out


def test_get_info_chunk_surgery_matches_reference(oracle, ref):
    """Chunk-level edits with VALID CRCs (bit flips almost always die at the first CRC check): inserted ancillary and unknown
    critical chunks, duplicated / missing / reordered / wrong-version fdEC, two IDATs, short IDAT, missing IEND, IHDR field edits
    with a recomputed CRC, chunk types outside A-Z/a-z, trailing bytes.  Product walk (host code of libfpng_b200.so) and the
    oracle's independently written walk must both return the unmodified reference's code (src/fpng.cpp:2930-3077)."""
    import fpng_b200
    n = 0
    seen = set()
    for png, w, h, c in files(oracle):
        sig, ch = _split(png)
        assert [t for t, _, _ in ch] == [b"IHDR", b"fdEC", b"IDAT", b"IEND"]
        ihdr, fdec, idat, iend = (r for _, _, r in ch)
        ihdr_body, idat_body = ch[0][1], ch[2][1]
        anc = _chunk(b"tEXt", b"Comment\0made by a test")
        cases = {
            "ancillary before fdEC": sig + ihdr + anc + fdec + idat + iend,
            "ancillary between fdEC and IDAT": sig + ihdr + fdec + anc + idat + iend,
            "ancillary after IDAT": sig + ihdr + fdec + idat + anc + iend,
            "empty ancillary": sig + ihdr + fdec + _chunk(b"zzZz", b"") + idat + iend,
            "private lower-case first letter, others upper": sig + ihdr + fdec + _chunk(b"aBCD", b"x") + idat + iend,
            "unknown critical chunk": sig + ihdr + _chunk(b"PLTE", bytes(6)) + fdec + idat + iend,
            "unknown critical after IDAT": sig + ihdr + fdec + idat + _chunk(b"ABCD", b"") + iend,
            "ancillary with bad crc": sig + ihdr + fdec + _chunk(b"tEXt", b"abc", good_crc=False) + idat + iend,
            "IEND with bad crc": sig + ihdr + fdec + idat + _chunk(b"IEND", b"", good_crc=False),
            "IDAT with bad crc (not checked by get_info)": sig + ihdr + fdec + _chunk(b"IDAT", idat_body, good_crc=False) + iend,
            "duplicate fdEC": sig + ihdr + fdec + fdec + idat + iend,
            "no fdEC": sig + ihdr + idat + iend,
            "fdEC after IDAT": sig + ihdr + idat + fdec + iend,
            "fdEC wrong version": sig + ihdr + _chunk(b"fdEC", bytes([82, 36, 147, 227, 1])) + idat + iend,
            "fdEC wrong signature": sig + ihdr + _chunk(b"fdEC", bytes([82, 36, 147, 226, 0])) + idat + iend,
            "fdEC too short": sig + ihdr + _chunk(b"fdEC", bytes([82, 36, 147, 227])) + idat + iend,
            "fdEC too long": sig + ihdr + _chunk(b"fdEC", bytes([82, 36, 147, 227, 0, 0])) + idat + iend,
            "two IDATs": sig + ihdr + fdec + idat + idat + iend,
            "IDAT split in two": sig + ihdr + fdec + _chunk(b"IDAT", idat_body[:9]) + _chunk(b"IDAT", idat_body[9:]) + iend,
            "IDAT of 6 bytes": sig + ihdr + fdec + _chunk(b"IDAT", idat_body[:6]) + iend,
            "IDAT of 7 bytes": sig + ihdr + fdec + _chunk(b"IDAT", idat_body[:7]) + iend,
            "no IDAT": sig + ihdr + fdec + iend,
            "no IEND": sig + ihdr + fdec + idat,
            "no IEND, ancillary last": sig + ihdr + fdec + idat + anc,
            "IEND first": sig + ihdr + iend + fdec + idat,
            "IEND with a body": sig + ihdr + fdec + idat + _chunk(b"IEND", b"tail"),
            "trailing bytes after IEND": png + b"\0" * 19,
            "chunk type with a digit": sig + ihdr + fdec + _chunk(b"tEX1", b"abc") + idat + iend,
            "chunk type with '[' (just above Z)": sig + ihdr + fdec + _chunk(b"tEX[", b"abc") + idat + iend,
            "chunk type with '`' (just below a)": sig + ihdr + fdec + _chunk(b"`EXt", b"abc") + idat + iend,
            "chunk type with '@' (just below A)": sig + ihdr + fdec + _chunk(b"@EXt", b"abc") + idat + iend,
            "chunk type with '{' (just above z)": sig + ihdr + fdec + _chunk(b"tEX{", b"abc") + idat + iend,
            "length field past the end": sig + ihdr + fdec + idat[:4] + b"IDAT",
            "length 0xFFFFFFFF": sig + ihdr + fdec + b"\xff\xff\xff\xff" + idat[4:] + iend,
            "second IHDR": sig + ihdr + ihdr + fdec + idat + iend,
            "IHDR length 14": sig + _chunk(b"IHDR", ihdr_body + b"\0") + fdec + idat + iend,
            "first chunk not IHDR but 13 bytes": sig + _chunk(b"tEXt", ihdr_body) + fdec + idat + iend,
        }
        def with_ihdr(**kw):
            b = bytearray(ihdr_body)
            for k, v in kw.items():
                if k == "w": b[0:4] = v.to_bytes(4, "big")
                elif k == "h": b[4:8] = v.to_bytes(4, "big")
                else: b[{"depth": 8, "ctype": 9, "comp": 10, "filt": 11, "lace": 12}[k]] = v
            return sig + _chunk(b"IHDR", bytes(b)) + fdec + idat + iend
        cases.update({
            "w = 0": with_ihdr(w=0), "h = 0": with_ihdr(h=0),
            "w = 2^24": with_ihdr(w=1 << 24, h=1), "w = 2^24 + 1": with_ihdr(w=(1 << 24) + 1, h=1), "h = 2^24 + 1": with_ihdr(w=1, h=(1 << 24) + 1),
            "w * h = 2^30": with_ihdr(w=1 << 15, h=1 << 15), "w * h = 2^30 + 2^15": with_ihdr(w=1 << 15, h=(1 << 15) + 1),
            "w = 0xFFFFFFFF": with_ihdr(w=0xFFFFFFFF),
            "depth 16": with_ihdr(depth=16), "depth 4": with_ihdr(depth=4),
            "grey": with_ihdr(ctype=0), "palette": with_ihdr(ctype=3), "grey + alpha": with_ihdr(ctype=4),
            "RGB <-> RGBA flipped": with_ihdr(ctype=2 if c == 4 else 6),
            "compression 1": with_ihdr(comp=1), "filter method 1": with_ihdr(filt=1), "interlaced": with_ihdr(lace=1),
            "bad depth and zero width (dimension check first)": with_ihdr(w=0, depth=1),
        })
        for name, bad in cases.items():
            exp = ref.get_info(bad)
            got = fpng_b200.fpng_get_info(bad)
            orc = oracle.get_info(bad)
            assert got[0] == exp[0], (name, got, exp)
            assert orc[0] == exp[0], (name, orc, exp)
            if exp[0] == 0:
                assert got[1:] == tuple(exp[1:4]) and tuple(orc[1:4]) == tuple(exp[1:4]), (name, got, orc, exp)
                st, ww, hh, cc, ofs, ln = fpng_b200.get_info_ex(bad)
                assert bad[ofs + 4:ofs + 8] == b"IDAT" and int.from_bytes(bad[ofs:ofs + 4], "big") == ln, name
            seen.add(exp[0])
            n += 1
    assert n > 350
    # the edits reach every code the walk can return (src/fpng.h:57-77): success, NOT_FPNG, NOT_PNG, HEADER_CRC32, INVALID_DIMENSIONS,
    # CHUNK_PARSING, INVALID_IDAT
    assert seen == {0, 1, 3, 4, 5, 7, 8}, seen


def test_get_info_random_chunk_sequences_match_reference(oracle, ref):
    """Property test: random chunk sequences drawn from a vocabulary of well-formed and damaged chunks (valid and invalid CRCs, bad type
    characters, fdEC / IDAT / IEND in any number and order, IHDR field variants, random truncation) -- product walk == oracle walk ==
    unmodified reference (code, and w/h/chans + IDAT location on success).  Deterministic (fixed seed)."""
    import fpng_b200
    rs = np.random.RandomState(2024)
    sig = bytes([137, 80, 78, 71, 13, 10, 26, 10])

    def ihdr():
        plain = rs.rand() < 0.7                                                 # mostly a valid header, so that the chunk rules are reached
        w = int(rs.choice([1, 7, 640, 65535, 65536, 1 << 15] if plain else [0, 1 << 24, (1 << 24) + 1, 1 << 15]))
        h = int(rs.choice([1, 9, 480, 1 << 15] if plain else [0, (1 << 15) + 1, 1 << 24, 3]))
        depth = 8 if plain else int(rs.choice([8, 16, 1])); ctype = int(rs.choice([2, 6] if plain else [2, 6, 0, 3, 4]))
        flags = [int(rs.rand() < (0.0 if plain else 0.1)) for _ in range(3)]
        body = w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([depth, ctype] + flags)
        if rs.rand() < 0.03:
            body += b"\0"
        return _chunk(b"IHDR", body, good_crc=rs.rand() > 0.03)

    def piece():
        k = rs.randint(0, 12)
        good = rs.rand() > 0.1
        if k <= 2:
            return _chunk(b"fdEC", bytes([82, 36, 147, 227, 0]) if rs.rand() > 0.15 else bytes(rs.randint(0, 256, rs.randint(0, 8), dtype=np.uint8)), good)
        if k <= 5:
            return _chunk(b"IDAT", bytes(rs.randint(0, 256, int(rs.choice([0, 6, 7, 8, 40, 200])), dtype=np.uint8)), good)
        if k <= 7:
            return _chunk(b"IEND", b"" if rs.rand() > 0.1 else b"x", good)
        if k == 8:
            return _chunk(bytes(rs.choice([b"tEXt", b"gAMA", b"zTXt", b"aaaa", b"prIv"])), bytes(rs.randint(0, 256, rs.randint(0, 20), dtype=np.uint8)), good)
        if k == 9:
            return _chunk(bytes(rs.choice([b"PLTE", b"ABCD", b"IDAU", b"Idat"])), bytes(rs.randint(0, 256, rs.randint(0, 20), dtype=np.uint8)), good)
        if k == 10:
            t = bytearray(b"tEXt"); t[rs.randint(0, 4)] = int(rs.choice([48, 64, 91, 96, 123, 0, 255]))
            return _chunk(bytes(t), b"abc", good)
        return bytes(rs.randint(0, 256, rs.randint(1, 16), dtype=np.uint8))      # raw garbage between chunks

    seen = {}
    for trial in range(3000):
        f = (sig if rs.rand() > 0.03 else sig[:7] + b"\x0b") + ihdr()
        if rs.rand() < 0.6:                                                     # bias towards the fpng layout so that deep rules are reached
            f += _chunk(b"fdEC", bytes([82, 36, 147, 227, 0]))
        for _ in range(rs.randint(1, 5)):
            f += piece()
        if rs.rand() < 0.7:
            f += _chunk(b"IEND", b"")
        if rs.rand() < 0.15:
            f = f[: rs.randint(0, len(f) + 1)]
        if not f:
            continue
        exp = ref.get_info(f)
        got = fpng_b200.fpng_get_info(f)
        orc = oracle.get_info(f)
        assert got[0] == exp[0] and orc[0] == exp[0], (trial, got, orc, exp, f[:80])
        if exp[0] == 0:
            assert got[1:] == tuple(exp[1:4]) and tuple(orc[1:4]) == tuple(exp[1:4])
            st, ww, hh, cc, ofs, ln = fpng_b200.get_info_ex(f)
            assert f[ofs + 4:ofs + 8] == b"IDAT" and int.from_bytes(f[ofs:ofs + 4], "big") == ln
        seen[exp[0]] = seen.get(exp[0], 0) + 1
    assert set(seen) == {0, 1, 3, 4, 5, 7, 8} and seen[0] >= 20, seen
