"""Developer smoke script for a gpurun call: encode a spread of shapes on the GPU and diff against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import imagegen
import fpng_b200
from oracle.pyoracle import Oracle, Ref

fpng_b200.fpng_init()
o = Oracle()
r = Ref() if Ref.available() else None
bad = 0
cases = []
for kind in ("g1", "g0", "g2", "runs", "zero", "mut"):
    for (w, h) in ((1, 1), (1, 7), (5, 3), (16, 16), (64, 9), (127, 5), (128, 4), (129, 11), (257, 33), (512, 512), (687, 101), (1000, 3)):
        for c in (3, 4):
            cases.append((kind, w, h, c))
for (kind, w, h, c) in cases:
    img = imagegen.make(kind, w, h, c, 7)
    for flags in (0, 1, 2):
        exp = o.encode(img, w, h, c, flags)
        ok, got = fpng_b200.fpng_encode_image_to_memory(img, w, h, c, flags)
        if not ok or got != exp:
            bad += 1
            if bad <= 25:
                n = min(len(got), len(exp))
                diff = next((i for i in range(n) if got[i] != exp[i]), n)
                print("MISMATCH", kind, w, h, c, "flags", flags, "len", len(got), len(exp), "first diff", diff)
print("cases", len(cases) * 3, "bad", bad)
