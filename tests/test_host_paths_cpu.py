"""CPU tests of the host-side code of libfpng_b200.so that runs without a device: the small host CRC-32 / Adler-32 routine the
container code needs (documented in include/fpng_b200.h: buffers < 4 KiB and any call before fpngb_init), and the argument /
state checks of the C ABI (reference: false / FPNG_DECODE_INVALID_ARG for bad arguments, src/fpng.cpp:1670-1680, 3092-3096).
No kernel can launch here (no GPU): every compute entry point must refuse loudly, never fall back."""
import ctypes as C
import zlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: these calls would initialise and run on the device")
    from fpng_b200._lib import lib
    lib_ = lib()
    lib_.fpngb_crc32.restype = C.c_uint32
    lib_.fpngb_crc32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    lib_.fpngb_adler32.restype = C.c_uint32
    lib_.fpngb_adler32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    lib_.fpngb_crc32_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
    lib_.fpngb_adler32_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
    return lib_


def test_host_checksums_known_answers_and_chaining(L, oracle):
    assert L.fpngb_is_initialized() == 0
    assert L.fpngb_crc32(b"123456789", 9, 0) == 0xCBF43926            # SURVEY 8c known answers
    assert L.fpngb_adler32(b"Wikipedia", 9, 1) == 0x11E60398
    assert L.fpngb_crc32(b"", 0, 0) == 0 and L.fpngb_adler32(b"", 0, 1) == 1
    rs = np.random.RandomState(3)
    for n in (1, 2, 3, 4, 5, 15, 16, 17, 63, 64, 65, 255, 256, 1000, 4095, 4096, 5552, 5553, 70000):
        buf = rs.randint(0, 256, n, dtype=np.uint8).tobytes()
        assert L.fpngb_crc32(buf, n, 0) == zlib.crc32(buf) == oracle.crc32(buf)
        assert L.fpngb_adler32(buf, n, 1) == zlib.adler32(buf) == oracle.adler32(buf)
        cut = n // 3                                                     # incremental chaining (src/fpng.h:27, 31)
        assert L.fpngb_crc32(buf[cut:], n - cut, L.fpngb_crc32(buf, cut, 0)) == zlib.crc32(buf)
        assert L.fpngb_adler32(buf[cut:], n - cut, L.fpngb_adler32(buf, cut, 1)) == zlib.adler32(buf)
        out = C.c_uint32(0)
        assert L.fpngb_crc32_ex(buf, n, 0, C.byref(out)) == 0 and out.value == zlib.crc32(buf)
        assert L.fpngb_adler32_ex(buf, n, 1, C.byref(out)) == 0 and out.value == zlib.adler32(buf)
    # worst case for the Adler sums: all 0xFF, longer than one 5552-byte reduction block
    buf = b"\xff" * 20000
    assert L.fpngb_adler32(buf, len(buf), 1) == zlib.adler32(buf)


def test_entry_points_refuse_without_a_device(L):
    ERR_INVALID_ARG, ERR_NOT_INITIALIZED, ERR_NO_DEVICE = 1, 3, 4
    DECODE_INVALID_ARG = 2
    assert L.fpngb_init(-1) == ERR_NO_DEVICE and L.fpngb_init(0) == ERR_NO_DEVICE
    assert L.fpngb_is_initialized() == 0
    px = np.zeros(4 * 4 * 4, np.uint8); out = np.zeros(4096, np.uint8); n = C.c_size_t(7)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.fpngb_encode_host(vp(px), 4, 4, 4, 0, vp(out), out.size, C.byref(n))
    assert rc in (ERR_NOT_INITIALIZED, ERR_NO_DEVICE) and not out.any()
    sizes = np.zeros(1, np.uint32)
    assert L.fpngb_encode_batch_host(vp(px), C.c_size_t(64), 1, 4, 4, 4, 0, vp(out), C.c_size_t(4096), vp(sizes)) in (ERR_NOT_INITIALIZED, ERR_NO_DEVICE)
    assert L.fpngb_encode_batch_device(vp(px), C.c_size_t(64), 1, 4, 4, 4, 0, vp(out), C.c_size_t(4096), vp(sizes), None) in (ERR_NOT_INITIALIZED, ERR_NO_DEVICE)
    # decode: the reference's INVALID_ARG cases first (null / zero size / bad desired channels, fpng.cpp:3092-3096) ...
    w, h, c = C.c_uint32(9), C.c_uint32(9), C.c_uint32(9)
    assert L.fpngb_decode_host(None, 10, vp(out), C.c_size_t(out.size), C.byref(w), C.byref(h), C.byref(c), 4) == DECODE_INVALID_ARG
    assert (w.value, h.value, c.value) == (0, 0, 0)                      # outputs are zeroed on entry (fpng.cpp:3087-3090)
    assert L.fpngb_decode_host(vp(px), 0, vp(out), C.c_size_t(out.size), C.byref(w), C.byref(h), C.byref(c), 4) == DECODE_INVALID_ARG
    assert L.fpngb_decode_host(vp(px), 64, vp(out), C.c_size_t(out.size), C.byref(w), C.byref(h), C.byref(c), 5) == DECODE_INVALID_ARG
    # ... then container errors are still reported by the host walk (FAILED_NOT_PNG = 3), never a decode on the CPU
    assert L.fpngb_decode_host(vp(px), 64, vp(out), C.c_size_t(out.size), C.byref(w), C.byref(h), C.byref(c), 4) == 3


def test_valid_file_is_not_decoded_on_the_cpu(L, oracle):
    """A well-formed fpng file passes the host container walk and then needs the device: without one the call must fail (the library
    reports FPNG_DECODE_INVALID_ARG for 'not initialised'), leaving the output untouched."""
    import imagegen
    w, h, c = 24, 5, 3
    png = oracle.encode(imagegen.make("g1", w, h, c, 1), w, h, c, 0)
    out = np.full(w * h * c, 0xAB, np.uint8)
    ww, hh, cc = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = L.fpngb_decode_host(png, len(png), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size), C.byref(ww), C.byref(hh), C.byref(cc), c)
    assert rc != 0 and (ww.value, hh.value, cc.value) == (w, h, c) and (out == 0xAB).all()
