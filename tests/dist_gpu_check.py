"""Multi-GPU check of the C-ABI gather (run under torchrun, world size >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py

Every rank encodes its own contiguous shard of a synthetic batch on its GPU, then the library's communicator
(fpngb_comm_init + fpngb_gather_setup + fpngb_gather_encoded_device: peer-window push over NVLink) gathers the encoded
files on rank 0, and -- second call -- on every rank.  Rank 0 (every rank, for the all-gather form) byte-compares EVERY
received file with the oracle's encoding of the corresponding image.  Prints one JSON line with the verdict and the
device-timed gather bandwidth."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import fpng_b200
    import imagegen
    from fpng_b200 import dist as fd
    from oracle.pyoracle import Oracle

    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fpng_b200.fpng_init(local)
    fd.init_comm()

    w, h, c = 640, 96, 4
    n_total = 8 * world + 3                       # uneven shards
    kinds = ["g1", "g0", "g2", "runs"]
    start, cnt = fd.shard_range(n_total, rank, world)
    nmax = max(fd.shard_range(n_total, r, world)[1] for r in range(world))
    imgs = np.stack([imagegen.make(kinds[i % 4], w, h, c, i) for i in range(start, start + cnt)])
    out, sizes = fpng_b200.encode_batch_device(torch.from_numpy(imgs).to(dev), 0)
    stride = int(out.stride(0))
    window_bytes = world * nmax * stride
    fd.gather_setup(window_bytes, nmax)
    ok_c, nr, rk, p2p = fd.comm_info()
    assert ok_c and nr == world and rk == rank

    o = Oracle()
    verdict = {"rank": rank, "p2p": p2p}
    stream = torch.cuda.current_stream(dev)
    for dst in (0, -1):
        ptrs = fd.gather_encoded_device(out, sizes, dst_rank=dst)
        win, offs, alls = fd.gathered_views(*ptrs, window_bytes, world, nmax, dev)
        torch.cuda.synchronize(dev)
        bad = 0
        if dst == -1 or rank == 0:
            offs_h = offs.cpu().numpy(); sz_h = alls.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            assert int(offs_h[-1]) >= 0 and int(offs_h[-1]) >> 63 == 0, "window overflow flagged"
            win_h = win[: int(offs_h[-1])].cpu().numpy()
            seen = 0
            for r in range(world):
                s0, cn = fd.shard_range(n_total, r, world)
                for i in range(nmax):
                    size = int(sz_h[r * nmax + i])
                    if i >= cn:
                        bad += size != 0
                        continue
                    exp = o.encode(imagegen.make(kinds[(s0 + i) % 4], w, h, c, s0 + i), w, h, c, 0)
                    got = win_h[int(offs_h[r * nmax + i]): int(offs_h[r * nmax + i]) + size].tobytes()
                    bad += got != exp
                    seen += 1
            bad += seen != n_total
        verdict[f"bad_dst{dst}"] = int(bad)
        dist.barrier()

    # bandwidth: a larger shard, device-timed, max over ranks
    w2, h2 = 1920, 1080
    big = torch.from_numpy(np.stack([imagegen.make("g1", w2, h2, 3, i + 100 * rank) for i in range(16)])).to(dev)
    out2, sizes2 = fpng_b200.encode_batch_device(big, 0)
    fd.gather_setup(world * 16 * int(out2.stride(0)), 16)
    for dst in (0, -1):
        fd.gather_encoded_device(out2, sizes2, dst_rank=dst)
        torch.cuda.synchronize(dev); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps = 5
        for _ in range(reps):
            fd.gather_encoded_device(out2, sizes2, dst_rank=dst)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        b = torch.tensor([float((sizes2.to(torch.int64) & 0xFFFFFFFF).sum().item())], dtype=torch.float64, device=dev)
        dist.all_reduce(b)
        mine = float((sizes2.to(torch.int64) & 0xFFFFFFFF).sum().item())
        verdict[f"gather_ms_dst{dst}"] = float(t.item())
        verdict[f"gbs_into_each_receiver_dst{dst}"] = (float(b.item()) - mine) / 1e9 / (float(t.item()) / 1e3)
    allv = [None] * world
    dist.all_gather_object(allv, verdict)
    fd.destroy_comm()
    if rank == 0:
        okay = all(v["bad_dst0"] == 0 and v["bad_dst-1"] == 0 for v in allv)
        print(json.dumps({"ok": okay, "world": world, "n_total": n_total, "ranks": allv}))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
