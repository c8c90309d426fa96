"""Multi-GPU plumbing: one process per GPU (torch.distributed), images sharded in contiguous blocks across ranks with
no collective on the data path, and ONE gather of the encoded buffers to the collecting rank afterwards
(BASELINE.json north_star).  NCCL over NVLink on the GPU box, gloo on CPU for the tests.

The reference has no distributed layer; this module is new.  It only moves bytes that the CUDA kernels produced.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, start+count) of images owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def compact(out, sizes):
    """[n, stride] files + [n] sizes -> (flat uint8 buffer with 16-byte aligned files, int64 offsets [n+1])."""
    import torch

    n = out.shape[0]
    if out.is_cuda:
        cap = int(out.shape[0]) * int(out.stride(0))
        dst = torch.empty((cap,), dtype=torch.uint8, device=out.device)
        offsets = torch.empty((n + 1,), dtype=torch.int64, device=out.device)
        s = torch.cuda.current_stream(out.device).cuda_stream
        check(lib().fpngb_compact_batch_device(out.data_ptr(), out.stride(0), sizes.data_ptr(), n, dst.data_ptr(), cap,
                                               offsets.data_ptr(), s), "compact")
        return dst, offsets
    # host tensors (gloo tests of the plumbing): same layout, plain slicing
    sz = sizes.to(torch.int64) & 0xFFFFFFFF
    padded = (sz + 15) // 16 * 16
    offsets = torch.zeros((n + 1,), dtype=torch.int64)
    offsets[1:] = torch.cumsum(padded, 0)
    dst = torch.zeros((int(offsets[-1]),), dtype=torch.uint8)
    for i in range(n):
        dst[int(offsets[i]): int(offsets[i]) + int(sz[i])] = out[i, : int(sz[i])]
    return dst, offsets


def gather_encoded(out, sizes, dst_rank: int = 0, group=None):
    """Gather every rank's encoded shard on `dst_rank`.

    out [n_local, stride] uint8, sizes [n_local] int32 (uint32 values).  All ranks must call.  Returns on dst_rank a list
    (one entry per rank, in rank order) of (flat_buffer, offsets[n_r+1], sizes[n_r]); None elsewhere.
    Exactly one size exchange (all_gather of padded size vectors) and one grouped send/recv of the compacted bytes.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = out.device
    n_local = out.shape[0]
    # 1. sizes: every rank learns n_r and the per-file sizes of every rank
    counts = [torch.zeros((1,), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    mine = torch.zeros((nmax,), dtype=torch.int32, device=dev)
    mine[:n_local] = sizes
    all_sizes = [torch.zeros((nmax,), dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, mine, group=group)
    # 2. bytes: compact locally, one grouped send/recv to the collecting rank
    flat, offsets = compact(out, sizes)
    total_local = int(offsets[-1].item())
    if rank != dst_rank:
        ops = [dist.P2POp(dist.isend, flat[:total_local], dst_rank, group)] if total_local else []
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None
    results, ops, bufs = [], [], {}
    for r in range(world):
        sz = (all_sizes[r][: counts[r]].to(torch.int64) & 0xFFFFFFFF)
        offs = torch.zeros((counts[r] + 1,), dtype=torch.int64, device=dev)
        offs[1:] = torch.cumsum((sz + 15) // 16 * 16, 0)
        tot = int(offs[-1].item())
        if r == rank:
            buf = flat[:total_local]
        else:
            buf = torch.empty((tot,), dtype=torch.uint8, device=dev)
            if tot:
                ops.append(dist.P2POp(dist.irecv, buf, r, group))
        results.append((buf, offs, sz))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return results


# ---------------------------------------------------------------------------------------------------------------------
# The C-ABI communicator (include/fpng_b200.h "multi-GPU"): NCCL communicator owned by the library + peer-window gather.
# torch.distributed is only the out-of-band channel that carries the 128-byte ncclUniqueId to every rank.
# ---------------------------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int, typestr: str, itemsize: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def init_comm(group=None) -> None:
    """Collective over `group`: creates the library's own NCCL communicator (fpngb_comm_init) on the fpng_init() device."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        check(lib().fpngb_comm_unique_id(idbuf.ctypes.data_as(C.c_void_p)), "comm_unique_id")
    t = torch.from_numpy(idbuf).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    idbuf = t.cpu().numpy().copy()
    check(lib().fpngb_comm_init(idbuf.ctypes.data_as(C.c_void_p), world, rank), "comm_init")


def destroy_comm() -> None:
    check(lib().fpngb_comm_destroy(), "comm_destroy")


def comm_info():
    n, r, p = C.c_int(), C.c_int(), C.c_int()
    rc = lib().fpngb_comm_info(C.byref(n), C.byref(r), C.byref(p))
    return rc == 0, n.value, r.value, bool(p.value)


def gather_setup(window_bytes: int, max_files_per_rank: int) -> None:
    """Collective: receive window on every rank + peer mappings (CUDA IPC over NVLink)."""
    check(lib().fpngb_gather_setup(window_bytes, max_files_per_rank), "gather_setup")


def gather_encoded_device(out, sizes, dst_rank: int = 0, stream=None):
    """Collective, stream-ordered, no host synchronisation on the peer-window path.  out [n_local, stride] uint8 CUDA,
    sizes [n_local] int32.  Returns zero-copy views (window uint8 [window_bytes], offsets int64 [nranks*nmax+1],
    all_sizes int32 [nranks*nmax]) of library-owned device memory; meaningful on the receiving rank(s) after `stream`."""
    import torch

    ok, nranks, rank, p2p = comm_info()
    assert ok, "init_comm() first"
    s = stream if stream is not None else torch.cuda.current_stream(out.device).cuda_stream
    win, offs, alls = C.c_void_p(), C.c_void_p(), C.c_void_p()
    n_local = int(out.shape[0])
    check(lib().fpngb_gather_encoded_device(out.data_ptr() if n_local else None, out.stride(0) if n_local else 16,
                                            sizes.data_ptr() if n_local else None, n_local, dst_rank,
                                            C.byref(win), C.byref(offs), C.byref(alls), s), "gather_encoded_device")
    return win.value, offs.value, alls.value


def gathered_views(win_ptr: int, offs_ptr: int, sizes_ptr: int, window_bytes: int, nranks: int, nmax: int, device):
    import torch

    window = torch.as_tensor(_DevArray(win_ptr, window_bytes, "|u1", 1), device=device)
    offsets = torch.as_tensor(_DevArray(offs_ptr, (nranks * nmax + 1) * 8, "<i8", 8), device=device)
    all_sizes = torch.as_tensor(_DevArray(sizes_ptr, nranks * nmax * 4, "<i4", 4), device=device)
    return window, offsets, all_sizes
