"""Small driver for ncu captures of the decode kernels: encodes a few workload images on the GPU and decodes them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import bench, fpng_b200

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = sys.argv[3] if len(sys.argv) > 3 else "g1"
fpng_b200.fpng_init(0)
dev = torch.device("cuda", 0)
batch = bench.make_device_batch(wl, kind, n, dev)
out, sizes = fpng_b200.encode_batch_device(batch, wl["flags"])
torch.cuda.synchronize()
sz = sizes.cpu().numpy().astype(np.int64)
files = [bytes(out[i, : sz[i]].cpu().numpy()) for i in range(n)]
t, stride, fs, ofs, lens, w, h, c = fpng_b200.pack_files_for_device(files, dev)
for _ in range(3):
    px, status = fpng_b200.decode_batch_device(t, fs, ofs, lens, w, h, c, c)
torch.cuda.synchronize()
print("ok", bool(torch.equal(px, batch)), int(status.abs().sum()))
