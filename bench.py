#!/usr/bin/env python
"""bench.py -- fpng hot-path benchmark (encode MP/s on synthetic batches) for the B200-native implementation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c1] [--kind g1|g0|g2]

One "step" = one pass of the encode hot path over one batch of synthetic images (device-resident for `value`,
host-resident pinned buffers through the C ABI for `e2e`).  Prints ONE JSON line on rank 0.  See DESIGN.md section
"Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import imagegen  # noqa: E402

MP = 1e6

# BASELINE.json configs.  `images` is the per-GPU batch (weak scaling: every rank encodes this many images).
WORKLOADS = {
    "c1": dict(name="C1 1 x 512x512 RGBA, 1-pass", w=512, h=512, chans=4, images=1, flags=0),
    "c2": dict(name="C2 256 x 1920x1080 RGB, 1-pass", w=1920, h=1080, chans=3, images=256, flags=0),
    "c3": dict(name="C3 128/GPU x 3840x2160 RGBA (green->alpha), 1-pass", w=3840, h=2160, chans=4, images=128, flags=0),
    "c4": dict(name="C4 128/GPU x 2048x2048 RGB, 2-pass (FPNG_ENCODE_SLOWER)", w=2048, h=2048, chans=3, images=128, flags=1),
}
N_UNIQUE_NOISE = 16


def workload_image(wl, kind: str, i: int) -> np.ndarray:
    """Image i of the synthetic batch (numpy twin of make_device_batch; used for the CPU arms and parity checks)."""
    w, h, c = wl["w"], wl["h"], wl["chans"]
    if kind == "g0":
        return imagegen.gradient(w, h, c, i)
    if kind == "g2":
        return imagegen.random_bytes(w, h, c, i % N_UNIQUE_NOISE)
    base = imagegen.gradient(w, h, c, i).astype(np.int16)
    rs = np.random.RandomState(1234 + (i % N_UNIQUE_NOISE))
    noise = rs.randint(-3, 4, size=base.shape).astype(np.int16)
    return ((base + noise) & 255).astype(np.uint8)


def make_device_batch(wl, kind: str, n: int, device, first_index: int = 0):
    """Builds the batch on the GPU: G0 gradient per image index + one of 16 host-generated noise fields."""
    import torch

    w, h, c = wl["w"], wl["h"], wl["chans"]
    x = torch.arange(w, device=device, dtype=torch.int32)[None, :]
    y = torch.arange(h, device=device, dtype=torch.int32)[:, None]
    r = ((255 * x) // max(w - 1, 1) + 0 * y).to(torch.int16)
    g = ((255 * y) // max(h - 1, 1) + 0 * x).to(torch.int16)
    batch = torch.empty((n, h, w, c), dtype=torch.uint8, device=device)
    noises = None
    if kind == "g1":
        u = min(N_UNIQUE_NOISE, n)
        host = np.stack([np.random.RandomState(1234 + j).randint(-3, 4, size=(h, w, c)).astype(np.int16) for j in range(u)])
        noises = torch.from_numpy(host).to(device)
    elif kind == "g2":
        u = min(N_UNIQUE_NOISE, n)
        host = np.stack([imagegen.random_bytes(w, h, c, j) for j in range(u)])
        noises = torch.from_numpy(host).to(device)
    for k in range(n):
        i = first_index + k
        if kind == "g2":
            batch[k] = noises[i % N_UNIQUE_NOISE]
            continue
        b = ((x + y + i) & 255).to(torch.int16)
        planes = [r, g, b] + ([g] if c == 4 else [])
        img = torch.stack(planes, dim=-1)
        if kind == "g1":
            img = img + noises[i % N_UNIQUE_NOISE]
        batch[k] = (img & 255).to(torch.uint8)
    return batch


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(wl, kind: str, seconds: float, steps: int = 1):
    """Times the UNMODIFIED reference encoder (oracle/_ref) on all host cores over a bounded sample of the workload.
    Returns (values MP/s per step, cores, sample description, kind)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.pyoracle import Oracle, Ref

    cores = os.cpu_count() or 1
    w, h, c, flags = wl["w"], wl["h"], wl["chans"], wl["flags"]
    use_ref = Ref.available()
    enc = Ref() if use_ref else Oracle()
    imgs = [np.ascontiguousarray(workload_image(wl, kind, i)) for i in range(min(cores, wl["images"]))]
    # calibrate: one image on one core
    t0 = time.perf_counter()
    (enc.encode_discard(imgs[0], w, h, c, flags, 1) if use_ref else enc.encode(imgs[0], w, h, c, flags))
    t1 = max(time.perf_counter() - t0, 1e-4)
    reps = max(1, int(seconds / t1))

    def work(k):
        img = imgs[k % len(imgs)]
        if use_ref:
            enc.encode_discard(img, w, h, c, flags, reps)
        else:
            for _ in range(reps):
                enc.encode(img, w, h, c, flags)

    vals = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for _ in range(steps):
            t0 = time.perf_counter()
            list(ex.map(work, range(cores)))
            dt = time.perf_counter() - t0
            vals.append(cores * reps * w * h / MP / dt)
    sample = f"{cores} threads x {reps} encodes of the first {len(imgs)} workload images ({kind}) per step, ~{seconds:.0f}s"
    return vals, cores, sample, ("reference" if use_ref else "port")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = max(args.steps + args.warmup, 1)
    per_step = max(1.0, min(8.0, 60.0 / total))
    vals, cores, sample, kind = cpu_reference_run(wl, args.kind, per_step, steps=args.warmup + args.steps)
    timed = vals[args.warmup:] or vals
    v = float(statistics.mean(timed))
    pixels_per_step = None
    line = {
        "impl": "reference", "metric": "encode_megapixels_per_sec", "value": v, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "kind": args.kind, "flags": wl["flags"], "where": "host CPU, reference fpng.cpp SSE4.1/PCLMUL build"},
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import fpng_b200
    from fpng_b200._lib import lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fpng_b200.fpng_init(local)
    L = lib()
    L.fpngb_profile_enable.argtypes = [C.c_int]
    L.fpngb_profile_read.argtypes = [C.POINTER(C.c_float), C.c_int]

    w, h, c, flags, n = wl["w"], wl["h"], wl["chans"], wl["flags"], args.images or wl["images"]
    batch = make_device_batch(wl, args.kind, n, dev, first_index=rank * n)
    stride = (fpng_b200.max_encoded_size(w, h, c) + 15) // 16 * 16
    out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    sizes = torch.empty((n,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        fpng_b200.encode_batch_device(batch, flags, out=out, sizes=sizes, stream=stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)

    # parity spot check against the reference/oracle on image 0 of this rank (outside the timed region)
    fsz = int(sizes[0].item()) & 0xFFFFFFFF
    got0 = bytes(out[0, :fsz].cpu().numpy())
    parity = None
    if rank == 0:
        from oracle.pyoracle import Oracle
        parity = got0 == Oracle().encode(workload_image(wl, args.kind, 0), w, h, c, flags)

    sampler = ClockSampler(local)
    L.fpngb_profile_enable(1)
    launches0 = fpng_b200.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = fpng_b200.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    prof = (C.c_float * 7)()
    L.fpngb_profile_read(prof, 7)
    L.fpngb_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_pixels = world * n * w * h * args.steps
    value = total_pixels / MP / (ms / 1e3)

    out_bytes = int((sizes.to(torch.int64) & 0xFFFFFFFF).sum().item())
    in_bytes = n * w * h * c
    names = ["hist", "huffman", "scan", "offsets", "pack", "adler", "crc"]
    kern = {k: float(v) for k, v in zip(names, prof)}
    peak, peak_src = measured_peak()
    algo = {"hist": in_bytes, "huffman": 0, "scan": in_bytes, "offsets": 0, "pack": in_bytes + out_bytes, "adler": 0, "crc": out_bytes}
    dominant = max(kern, key=lambda k: kern[k])

    def roof(k):
        gbs = algo[k] / 1e9 / (kern[k] / 1e3) if kern[k] > 0 else 0.0
        return {"kernel": k, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                "ms_per_launch": kern[k], "algorithmic_bytes_per_launch": algo[k], "peak_source": peak_src}

    line = {
        "metric": "encode_megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "kind": args.kind, "images_per_gpu": n, "w": w, "h": h, "chans": c, "flags": flags,
                   "l2": "inputs larger than L2 (%.0f MB per step)" % (in_bytes / 1e6) if in_bytes > 126e6 else "input smaller than L2: L2-warm",
                   "out_over_in": out_bytes / in_bytes, "parity_image0_vs_oracle": parity},
        "clocks": clocks, "gpu_launches": int(launches),
        "kernels_ms": kern, "roofline": roof(dominant), "roofline_scan": roof("scan"),
        "whole_step": {"algorithmic_gbs": (in_bytes + out_bytes) / 1e9 / (ms / args.steps / 1e3),
                       "frac_of_peak": (in_bytes + out_bytes) / 1e9 / (ms / args.steps / 1e3) / peak},
    }

    # ---- end to end through the C ABI with host (pinned) buffers: H2D + kernels + D2H inside the timed region
    n_e2e = min(n, args.e2e_images) if args.e2e_images else n
    hin = torch.empty((n_e2e, h, w, c), dtype=torch.uint8).pin_memory()
    hin.copy_(batch[:n_e2e].cpu())
    hout = torch.empty((n_e2e, stride), dtype=torch.uint8).pin_memory()
    hsizes = np.zeros(n_e2e, dtype=np.uint32)

    def e2e_step():
        rc = L.fpngb_encode_batch_host(hin.data_ptr(), h * w * c, n_e2e, w, h, c, flags, hout.data_ptr(), stride, hsizes.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(f"fpngb_encode_batch_host failed: {rc}")

    e2e_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    e2e_ok = bytes(hout[0, : int(hsizes[0])].numpy()) == got0
    line["e2e"] = {"value": world * n_e2e * w * h * e2e_steps / MP / dt, "unit": "MP/s", "h2d_bytes_per_step": n_e2e * w * h * c,
                   "d2h_bytes_per_step": int(hsizes.astype(np.int64).sum()) + 4 * n_e2e, "images_per_step": n_e2e, "steps": e2e_steps,
                   "api": "fpngb_encode_batch_host (C ABI, pinned host buffers, blocking)", "timer": "host wall clock around the blocking calls",
                   "matches_device_path": bool(e2e_ok)}

    if rank == 0 and world == 1 and not args.no_cpu:
        vals, cores, sample, kind = cpu_reference_run(wl, args.kind, args.cpu_seconds)
        line["cpu_baseline"] = {"value": float(vals[0]), "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--kind", default="g1", choices=["g0", "g1", "g2"])
    ap.add_argument("--images", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--e2e-images", type=int, default=0, help="images per e2e step (0 = the whole per-GPU batch)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
