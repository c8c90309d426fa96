"""fpng_b200.decode_files (batch decode of files with MIXED dimensions: host container walk, grouping by shape, one
fpngb_decode_batch_host call per group) checked on the CPU: the container walk is the library's real host code; only the device call
is replaced by a stand-in that decodes with the oracle, so the grouping / ordering / status plumbing is what is tested here.  The
device call itself is covered by tests/test_decode_gpu.py::test_decode_batch_host."""
import ctypes as C

import numpy as np
import pytest

import imagegen


class _LibWithOracleDecode:
    """the real library for every symbol except fpngb_decode_batch_host"""

    def __init__(self, real, oracle):
        self._real, self._oracle, self.calls = real, oracle, []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def fpngb_decode_batch_host(self, ptrs, sizes_p, n, desired, optr, out_stride, w_p, h_p, c_p, status_p):
        sizes = np.ctypeslib.as_array(C.cast(sizes_p, C.POINTER(C.c_uint32)), (n,))
        status = np.ctypeslib.as_array(C.cast(status_p, C.POINTER(C.c_int32)), (n,))
        shapes = set()
        for i in range(n):
            data = C.string_at(ptrs[i], int(sizes[i]))
            st, px, w, h, c = self._oracle.decode(data, desired)
            status[i] = st
            if st == 0:
                shapes.add((w, h, c))
                C.memmove(optr + i * out_stride, px.ctypes.data, px.size)
        assert len(shapes) <= 1, "one shape per C-ABI call"
        self.calls.append((n, shapes))
        if shapes:
            (w, h, c), = shapes
            w_p._obj.value, h_p._obj.value, c_p._obj.value = w, h, c
        return 0


def test_decode_files_groups_mixed_shapes(oracle, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the GPU suite")
    import fpng_b200
    from fpng_b200 import api, _lib
    fake = _LibWithOracleDecode(_lib.lib(), oracle)
    monkeypatch.setattr(api, "lib", lambda: fake)
    specs = [("g1", 40, 9, 3, 0), ("g0", 64, 16, 4, 0), ("g1", 40, 9, 3, 1), ("runs", 64, 16, 4, 0), ("g2", 7, 3, 3, 0), ("g1", 40, 9, 4, 2),
             ("g1", 64, 16, 4, 1)]
    imgs = [np.asarray(imagegen.make(k, w, h, c, i)).reshape(h, w, c) for i, (k, w, h, c, f) in enumerate(specs)]
    files = [oracle.encode(im, s[1], s[2], s[3], s[4]) for im, s in zip(imgs, specs)]
    bad_crc = bytearray(files[0]); bad_crc[20] ^= 1                       # IHDR damage: FPNG_DECODE_FAILED_HEADER_CRC32 from the host walk
    corrupt = bytearray(files[1]); corrupt[80] ^= 0xFF                    # stream damage: FPNG_DECODE_NOT_FPNG (or still decodable) from the decoder
    batch = files[:3] + [bytes(bad_crc), b"", files[3], b"not a png", bytes(corrupt)] + files[4:]
    for desired in (3, 4):
        fake.calls.clear()
        res = fpng_b200.decode_files(batch, desired)
        assert len(res) == len(batch)
        for f, (st, px, w, h, c) in zip(batch, res):
            if len(f) == 0:
                assert st == fpng_b200.FPNG_DECODE_INVALID_ARG and px is None
                continue
            est, epx, ew, eh, ec = oracle.decode(f, desired)
            assert st == est, (st, est)
            if est == 0:
                assert (w, h, c) == (ew, eh, ec) and np.array_equal(px, epx)
            else:
                assert px is None
        # four distinct shapes among the decodable files -> four device calls, each with a single shape
        assert len(fake.calls) == 4 and sorted(n for n, _ in fake.calls) == [1, 1, 2, 4]
    assert [r[0] for r in fpng_b200.decode_files(batch[:2], 5)] == [fpng_b200.FPNG_DECODE_INVALID_ARG] * 2
