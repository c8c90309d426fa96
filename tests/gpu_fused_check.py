"""Dev check of the single-pass encoder against the oracle over many shapes (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
import imagegen, fpng_b200
from oracle.pyoracle import Oracle
o = Oracle(); fpng_b200.fpng_init()
from fpng_b200._lib import lib
L = lib()
mode = sys.argv[1] if len(sys.argv) > 1 else "inline"
L.fpngb_debug_inline_crc(1 if mode == "inline" else 0)
L.fpngb_debug_use_fused(0 if mode in ("old", "old_serial") else 1)
L.fpngb_debug_crc_overlap(0 if mode == "old_serial" else 1)
print("mode", mode)
bad = 0; n = 0
shapes = [(16, 1), (16, 2), (32, 3), (512, 9), (528, 17), (1024, 5), (1040, 33), (1920, 8), (2048, 4), (4096, 3), (4080, 7), (3840, 5), (64, 300), (1536, 11),
          (1, 1), (5, 3), (85, 2), (687, 41), (513, 6), (2049, 3), (4095, 2), (341, 25)]
for kind in ('g1', 'g0', 'runs', 'g2', 'mut', 'zero'):
    for (w, h) in shapes:
        for c in (3, 4):
            for flags in (0, 1):
                imgs = np.stack([imagegen.make(kind, w, h, c, 3 + i) for i in range(3)])
                out, sizes = fpng_b200.encode_batch_device(torch.from_numpy(imgs).cuda(), flags)
                torch.cuda.synchronize()
                sz = sizes.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
                oh = out.cpu().numpy()
                for i in range(3):
                    exp = o.encode(imgs[i], w, h, c, flags)
                    got = oh[i, :sz[i]].tobytes()
                    n += 1
                    if got != exp:
                        bad += 1
                        if bad < 12:
                            first = next((k for k in range(min(len(got), len(exp))) if got[k] != exp[k]), -1)
                            print("MISMATCH", kind, w, h, c, flags, i, len(got), len(exp), "first diff byte", first)
print("checked", n, "bad", bad)
sys.exit(1 if bad else 0)
