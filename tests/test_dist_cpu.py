"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU plumbing in fpng_b200/dist.py: sharding covers every image
exactly once, and the single gather of variable-size encoded buffers delivers every rank's files intact to rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fpng_b200.dist import compact, gather_encoded, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 9, 256, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(n, r, world)
                seen.extend(range(s, s + c))
            assert seen == list(range(n))
            counts = [shard_range(n, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _fake_shard(rank, n_local, stride):
    rs = np.random.RandomState(100 + rank)
    sizes = rs.randint(74, stride, size=n_local).astype(np.int32)
    out = rs.randint(0, 256, size=(n_local, stride), dtype=np.uint8)
    return torch.from_numpy(out), torch.from_numpy(sizes)


def test_compact_layout():
    out, sizes = _fake_shard(0, 5, 300)
    flat, offs = compact(out, sizes)
    assert offs[0] == 0 and all(int(o) % 16 == 0 for o in offs)
    for i in range(5):
        assert torch.equal(flat[int(offs[i]): int(offs[i]) + int(sizes[i])], out[i, : int(sizes[i])])


def _worker(rank, world, port, n_total, stride, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, cnt = shard_range(n_total, rank, world)
        out, sizes = _fake_shard(rank, cnt, stride)
        res = gather_encoded(out, sizes, dst_rank=0)
        ok = True
        if rank == 0:
            assert res is not None and len(res) == world
            for r in range(world):
                _, c = shard_range(n_total, r, world)
                eout, esizes = _fake_shard(r, c, stride)
                buf, offs, sz = res[r]
                ok &= len(sz) == c
                for i in range(c):
                    ok &= int(sz[i]) == int(esizes[i])
                    ok &= bool(torch.equal(buf[int(offs[i]): int(offs[i]) + int(sz[i])], eout[i, : int(esizes[i])]))
        else:
            ok &= res is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 7), (3, 10), (2, 1)])
def test_gather_encoded_gloo(world, n_total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, 400, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
