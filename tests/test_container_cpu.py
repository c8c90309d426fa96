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
