"""Python mirror of the reference's operator interface (src/fpng.h:13-111) over the C ABI.

Reference signature                                        -> here
  fpng_init()                                              -> fpng_init(device=-1)
  fpng_encode_image_to_memory(pImage,w,h,chans,out,flags)  -> fpng_encode_image_to_memory(image,w,h,chans,flags) -> (ok, bytes)
  fpng_decode_memory(pImage,size,out,w,h,chans,desired)    -> fpng_decode_memory(data, desired) -> (status, pixels, w, h, chans)
Errors follow the reference: encode returns (False, b"") on invalid arguments, decode returns the FPNG_DECODE_* codes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import FpngB200Error, check, lib

FPNG_ENCODE_SLOWER = 1          # src/fpng.h:38
FPNG_FORCE_UNCOMPRESSED = 2     # src/fpng.h:41
(FPNG_DECODE_SUCCESS, FPNG_DECODE_NOT_FPNG, FPNG_DECODE_INVALID_ARG, FPNG_DECODE_FAILED_NOT_PNG,
 FPNG_DECODE_FAILED_HEADER_CRC32, FPNG_DECODE_FAILED_INVALID_DIMENSIONS, FPNG_DECODE_FAILED_DIMENSIONS_TOO_LARGE,
 FPNG_DECODE_FAILED_CHUNK_PARSING, FPNG_DECODE_FAILED_INVALID_IDAT, FPNG_DECODE_FILE_OPEN_FAILED,
 FPNG_DECODE_FILE_TOO_LARGE, FPNG_DECODE_FILE_READ_FAILED, FPNG_DECODE_FILE_SEEK_FAILED) = range(13)   # src/fpng.h:57-77


def _u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def fpng_init(device: int = -1) -> None:
    """src/fpng.h:17. Brings up the CUDA side (device, stream, code books). Raises when no GPU is present."""
    check(lib().fpngb_init(device), "fpng_init")


def fpng_cpu_supports_sse41() -> bool:
    """src/fpng.h:23. Kept for API completeness; this implementation has no CPU SIMD path, so always False."""
    if not lib().fpngb_is_initialized():
        raise FpngB200Error("fpng_cpu_supports_sse41: fpng_init() must be called first")
    return False


def launch_count() -> int:
    return int(lib().fpngb_launch_count())


def max_encoded_size(w: int, h: int, chans: int) -> int:
    return int(lib().fpngb_max_encoded_size(w, h, chans))


def fpng_crc32(data, prev_crc32: int = 0) -> int:
    a = _u8(data)
    return int(lib().fpngb_crc32(a.ctypes.data_as(C.c_void_p), a.size, prev_crc32))


def fpng_adler32(data, adler: int = 1) -> int:
    a = _u8(data)
    return int(lib().fpngb_adler32(a.ctypes.data_as(C.c_void_p), a.size, adler))


def fpng_encode_image_to_memory(image, w: int, h: int, num_chans: int, flags: int = 0):
    """src/fpng.h:48 / src/fpng.cpp:1662.  Returns (ok, png_bytes)."""
    L = lib()
    if w < 1 or h < 1 or num_chans not in (3, 4) or w > (1 << 24) or h > (1 << 24) or w * h > 0xFFFFFFFF:
        return False, b""
    a = _u8(image)
    if a.size != w * h * num_chans:
        return False, b""
    cap = max_encoded_size(w, h, num_chans)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = L.fpngb_encode_host(a.ctypes.data_as(C.c_void_p), w, h, num_chans, flags, out.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    if rc == 1:
        return False, b""
    check(rc, "fpng_encode_image_to_memory")
    return True, out[: n.value].tobytes()


def fpng_encode_image_to_file(filename: str, image, w: int, h: int, num_chans: int, flags: int = 0) -> bool:
    """src/fpng.h:52."""
    ok, data = fpng_encode_image_to_memory(image, w, h, num_chans, flags)
    if not ok:
        return False
    try:
        with open(filename, "wb") as f:
            f.write(data)
    except OSError:
        return False
    return True


def fpng_get_info(data):
    """src/fpng.h:92. Returns (status, w, h, channels_in_file)."""
    a = _u8(data)
    w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    st = lib().fpngb_get_info(a.ctypes.data_as(C.c_void_p), a.size, C.byref(w), C.byref(h), C.byref(c))
    return st, w.value, h.value, c.value


def fpng_decode_memory(data, desired_channels: int):
    """src/fpng.h:108. Returns (status, pixels|None, w, h, channels_in_file)."""
    a = _u8(data)
    if a.size == 0 or desired_channels not in (3, 4):
        return FPNG_DECODE_INVALID_ARG, None, 0, 0, 0
    st, w, h, c = fpng_get_info(a)
    if st != FPNG_DECODE_SUCCESS:
        return st, None, w, h, c
    need = w * h * desired_channels
    if need > 0xFFFFFFFF:
        return FPNG_DECODE_FAILED_DIMENSIONS_TOO_LARGE, None, w, h, c
    out = np.empty(need, dtype=np.uint8)
    ww, hh, cc = C.c_uint32(), C.c_uint32(), C.c_uint32()
    st = lib().fpngb_decode_host(a.ctypes.data_as(C.c_void_p), a.size, out.ctypes.data_as(C.c_void_p), need,
                                 C.byref(ww), C.byref(hh), C.byref(cc), desired_channels)
    return st, (out if st == 0 else None), ww.value, hh.value, cc.value


def fpng_decode_file(filename: str, desired_channels: int):
    """src/fpng.h:111."""
    try:
        with open(filename, "rb") as f:
            data = f.read()
    except OSError:
        return FPNG_DECODE_FILE_OPEN_FAILED, None, 0, 0, 0
    if len(data) > 0xFFFFFFFF:
        return FPNG_DECODE_FILE_TOO_LARGE, None, 0, 0, 0
    return fpng_decode_memory(data, desired_channels)


def encode_batch_device(images, flags: int = 0, out=None, sizes=None, stream=None):
    """Encode a batch of equally sized device-resident images.

    images: torch.uint8 CUDA tensor [n, h, w, chans] (contiguous). Returns (out [n, stride] uint8, sizes [n] int32 holding
    uint32 file sizes) on the same device; kernels are enqueued on the current torch stream (or `stream`), not synchronised.
    """
    import torch

    assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous() and images.dim() == 4
    n, h, w, c = images.shape
    stride = (max_encoded_size(w, h, c) + 15) // 16 * 16
    if out is None:
        out = torch.empty((n, stride), dtype=torch.uint8, device=images.device)
    if sizes is None:
        sizes = torch.empty((n,), dtype=torch.int32, device=images.device)
    s = stream if stream is not None else torch.cuda.current_stream(images.device).cuda_stream
    rc = lib().fpngb_encode_batch_device(images.data_ptr(), h * w * c, n, w, h, c, flags, out.data_ptr(), out.stride(0),
                                         sizes.data_ptr(), s)
    check(rc, "encode_batch_device")
    return out, sizes


def get_info_ex(data):
    """fpng_get_info plus (idat_ofs, idat_len): the host-side container walk (src/fpng.cpp:2930-3077)."""
    a = _u8(data)
    w, h, c, o, l = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    st = lib().fpngb_get_info_ex(a.ctypes.data_as(C.c_void_p), a.size, C.byref(w), C.byref(h), C.byref(c), C.byref(o), C.byref(l))
    return st, w.value, h.value, c.value, o.value, l.value


def pack_files_for_device(files, device=None):
    """Container-walks a list of fpng files (bytes) of identical dimensions on the host and packs them into one device
    tensor [n, stride].  Returns (tensor, stride, sizes, idat_ofs, idat_len, w, h, chans)."""
    import torch

    infos = [get_info_ex(f) for f in files]
    for st, *_ in infos:
        if st != 0:
            raise FpngB200Error(f"pack_files_for_device: fpng_get_info returned {st}")
    w, h, c = infos[0][1:4]
    if any(i[1:4] != (w, h, c) for i in infos):
        raise FpngB200Error("pack_files_for_device: files differ in dimensions/channels")
    stride = (max(len(f) for f in files) + 16 + 15) // 16 * 16
    host = np.zeros((len(files), stride), dtype=np.uint8)
    for i, f in enumerate(files):
        host[i, : len(f)] = np.frombuffer(f, dtype=np.uint8)
    t = torch.from_numpy(host)
    if device is not None:
        t = t.to(device)
    sizes = np.array([len(f) for f in files], dtype=np.uint32)
    ofs = np.array([i[4] for i in infos], dtype=np.uint32)
    lens = np.array([i[5] for i in infos], dtype=np.uint32)
    return t, stride, sizes, ofs, lens, w, h, c


def decode_batch_device(files_dev, sizes, idat_ofs, idat_len, w: int, h: int, chans_in_file: int, desired_channels: int,
                        out=None, status=None, stream=None):
    """Decode n fpng files resident on the device (tensor [n, stride] uint8). Returns (pixels [n, h, w, desired], status [n])."""
    import torch

    assert files_dev.is_cuda and files_dev.dtype == torch.uint8 and files_dev.dim() == 2 and files_dev.is_contiguous()
    n = files_dev.shape[0]
    if out is None:
        out = torch.empty((n, h, w, desired_channels), dtype=torch.uint8, device=files_dev.device)
    if status is None:
        status = torch.empty((n,), dtype=torch.int32, device=files_dev.device)
    s = stream if stream is not None else torch.cuda.current_stream(files_dev.device).cuda_stream
    sizes = np.ascontiguousarray(sizes, dtype=np.uint32); idat_ofs = np.ascontiguousarray(idat_ofs, dtype=np.uint32)
    idat_len = np.ascontiguousarray(idat_len, dtype=np.uint32)
    rc = lib().fpngb_decode_batch_device(files_dev.data_ptr(), files_dev.stride(0), sizes.ctypes.data_as(C.c_void_p),
                                         idat_ofs.ctypes.data_as(C.c_void_p), idat_len.ctypes.data_as(C.c_void_p), n, w, h,
                                         chans_in_file, desired_channels, out.data_ptr(), h * w * desired_channels,
                                         status.data_ptr(), s)
    check(rc, "decode_batch_device")
    return out, status


def decode_batch_host(file_ptrs, sizes, desired_channels: int, out, out_stride: int):
    """Pipelined host-buffer batch decode (fpngb_decode_batch_host).  file_ptrs: array of host addresses (uint64), sizes:
    uint32 array, out: host buffer (numpy uint8 / pinned torch tensor) with out_stride bytes per image.
    Returns (rc, w, h, chans, status[n])."""
    n = len(sizes)
    ptrs = (C.c_void_p * n)(*[int(p) for p in file_ptrs])
    sz = np.ascontiguousarray(sizes, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    optr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
    rc = lib().fpngb_decode_batch_host(ptrs, sz.ctypes.data_as(C.c_void_p), n, desired_channels, optr, out_stride,
                                       C.byref(w), C.byref(h), C.byref(c), status.ctypes.data_as(C.c_void_p))
    return rc, w.value, h.value, c.value, status


def decode_files(files, desired_channels: int):
    """Batch form of fpng_decode_memory (src/fpng.h:108) for a list of files (bytes-like) of ANY mix of dimensions: the C ABI decodes
    one shape per call (one kernel grid per shape), so the files are container-walked on the host, grouped by (width, height,
    channels) and every group goes through fpngb_decode_batch_host.  Returns one (status, pixels or None, w, h, chans_in_file) per
    file, in input order; status is the FPNG_DECODE_* code fpng_decode_memory would return for that file."""
    if desired_channels not in (3, 4):
        return [(FPNG_DECODE_INVALID_ARG, None, 0, 0, 0) for _ in files]
    bufs = [_u8(f) for f in files]
    results = [None] * len(bufs)
    groups = {}
    for i, a in enumerate(bufs):
        if a.size == 0:
            results[i] = (FPNG_DECODE_INVALID_ARG, None, 0, 0, 0)
            continue
        st, w, h, c, _, _ = get_info_ex(a)
        if st != FPNG_DECODE_SUCCESS:
            results[i] = (st, None, w, h, c)
        elif w * h * desired_channels > 0xFFFFFFFF:
            results[i] = (FPNG_DECODE_FAILED_DIMENSIONS_TOO_LARGE, None, w, h, c)                 # fpng.cpp:3103-3105
        else:
            groups.setdefault((w, h, c), []).append(i)
    for (w, h, c), idx in groups.items():
        need = w * h * desired_channels
        out = np.empty((len(idx), need), dtype=np.uint8)
        rc, ww, hh, cc, status = decode_batch_host([bufs[i].ctypes.data for i in idx], [bufs[i].size for i in idx], desired_channels, out, need)
        check(rc, "decode_files")
        for k, i in enumerate(idx):
            ok = int(status[k]) == FPNG_DECODE_SUCCESS
            results[i] = (int(status[k]), out[k].copy() if ok else None, w, h, c)
    return results


# ---- static-table training (reference: FPNG_TRAIN_HUFFMAN_TABLES / fpng_test -t; src/fpng.h:114-120) ----
def train_accumulate_device(images, counts=None, stream=None):
    """Adds the 16-bit scaled symbol counts of every image of a CUDA uint8 batch [n, h, w, chans] to counts[288] (uint64)."""
    import torch

    assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous() and images.dim() == 4
    n, h, w, c = images.shape
    if counts is None:
        counts = np.zeros(288, dtype=np.uint64)
    s = stream if stream is not None else torch.cuda.current_stream(images.device).cuda_stream
    check(lib().fpngb_train_accumulate_device(images.data_ptr(), h * w * c, n, w, h, c, counts.ctypes.data_as(C.c_void_p), s), "train_accumulate_device")
    return counts


def create_dynamic_block_prefix(counts, num_chans: int):
    """src/fpng.cpp:910. Returns (prefix bytes, bit_buf, bit_buf_size, codes[288], code sizes[288])."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    prefix = np.zeros(4096, np.uint8); n = C.c_size_t(); bb = C.c_uint64(); bs = C.c_int()
    codes = np.zeros(288, np.uint32); sizes = np.zeros(288, np.uint8)
    check(lib().fpngb_create_dynamic_block_prefix(counts.ctypes.data_as(C.c_void_p), num_chans, prefix.ctypes.data_as(C.c_void_p), prefix.size,
                                                  C.byref(n), C.byref(bb), C.byref(bs), codes.ctypes.data_as(C.c_void_p),
                                                  sizes.ctypes.data_as(C.c_void_p)), "create_dynamic_block_prefix")
    return prefix[: n.value].tobytes(), bb.value, bs.value, codes, sizes


def set_static_table(num_chans: int, prefix: bytes = b"", bit_buf: int = 0, bit_buf_size: int = 0) -> None:
    """Install a trained 1-pass table for num_chans (empty prefix restores the built-in table)."""
    a = np.frombuffer(prefix, dtype=np.uint8) if prefix else np.zeros(1, np.uint8)
    check(lib().fpngb_set_static_table(num_chans, a.ctypes.data_as(C.c_void_p), len(prefix), bit_buf, bit_buf_size), "set_static_table")
