"""Small encode+decode workload for compute-sanitizer (memcheck / racecheck / synccheck) runs on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import imagegen, fpng_b200
from oracle.pyoracle import Oracle

o = Oracle(); fpng_b200.fpng_init(0)
bad = 0
cases = [("g1", 528, 5, 3), ("g1", 516, 4, 4), ("g0", 1024, 3, 3), ("runs", 640, 4, 4), ("g2", 64, 6, 3), ("mut", 333, 7, 3), ("g1", 37, 9, 4), ("zero", 1040, 2, 4)]
from fpng_b200 import _lib
for rpw in (5, 1):   # both shapes of the pack kernel: 5 scanlines per warp (large batches) and 1 (small jobs)
    _lib.lib().fpngb_debug_rows_per_warp(rpw)
    for kind, w, h, c in cases:
        img = imagegen.make(kind, w, h, c, 5)
        for flags in (0, 1, 2):
            ok, png = fpng_b200.fpng_encode_image_to_memory(img, w, h, c, flags)
            bad += (not ok) or png != o.encode(img, w, h, c, flags)
            if rpw == 5:
                for d in (3, 4):
                    st, px, *_ = fpng_b200.fpng_decode_memory(png, d)
                    bad += st != 0 or not np.array_equal(px, o.decode(png, d)[1])
_lib.lib().fpngb_debug_rows_per_warp(0)
# the opt-in staged write pass of the decoder (and the per-thread one after it) on the same files
_lib.lib().fpngb_debug_decode_staged(1)
for kind, w, h, c in cases:
    img = imagegen.make(kind, w, h, c, 5)
    png = o.encode(img, w, h, c, 0)
    for d in (3, 4):
        st, px, *_ = fpng_b200.fpng_decode_memory(png, d)
        bad += st != 0 or not np.array_equal(px, o.decode(png, d)[1])
_lib.lib().fpngb_debug_decode_staged(0)
d = np.random.RandomState(0).randint(0, 256, 70000, dtype=np.uint8)
bad += fpng_b200.fpng_crc32(d) != o.crc32(d)
bad += fpng_b200.fpng_adler32(d) != o.adler32(d)
print("sanitize_driver bad =", bad)
sys.exit(1 if bad else 0)
