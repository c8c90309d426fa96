"""Generates tests/golden/reference_vectors.json from the UNMODIFIED reference encoder/decoder (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run here (the reference tree is not available on the GPU box):

    python tests/golden/make_golden.py

Each vector: generator parameters (tests/imagegen.py is deterministic), encode flags, sha256 + size of the reference's
PNG bytes, sha256 of the pixels (what every decoder must return), and for tiny images the full PNG bytes (hex).
The reference ships no golden vectors of its own, so these are the pinned known answers.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import imagegen  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

CASES = []
for kind in ("g0", "g1", "g2", "runs", "mut", "zero"):
    for (w, h) in ((1, 1), (2, 1), (1, 5), (7, 3), (16, 1), (24, 2), (85, 4), (86, 3), (127, 9), (128, 8), (129, 8),
                   (255, 5), (256, 4), (340, 7), (513, 17), (687, 23), (1024, 6)):
        for c in (3, 4):
            CASES.append((kind, w, h, c))
CASES += [("g1", 1920, 64, 3), ("g1", 3840, 24, 4), ("g0", 1920, 64, 3), ("g0", 3840, 24, 4), ("g2", 2048, 16, 3),
          ("runs", 4096, 8, 4), ("runs", 2049, 8, 3)]


def main():
    r = Ref()
    vectors = []
    for idx, (kind, w, h, c) in enumerate(CASES):
        img = imagegen.make(kind, w, h, c, idx)
        pix = img.tobytes()
        for flags in (0, 1, 2):
            png = r.encode(img, w, h, c, flags)
            st, px, ww, hh, cc = r.decode(png, c)
            assert st == 0 and px.tobytes() == pix
            v = {"kind": kind, "w": w, "h": h, "chans": c, "index": idx, "flags": flags, "size": len(png),
                 "png_sha256": hashlib.sha256(png).hexdigest(), "pixels_sha256": hashlib.sha256(pix).hexdigest()}
            if len(png) <= 160:
                v["png_hex"] = png.hex()
            vectors.append(v)
    out = {"reference": "richgel999/fpng v1.0.6 (src/fpng.cpp), g++ -O3 -msse4.1 -mpclmul -DFPNG_NO_SSE=0",
           "generator": "tests/imagegen.py make(kind, w, h, chans, index)", "vectors": vectors}
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(vectors), "vectors")


if __name__ == "__main__":
    main()
