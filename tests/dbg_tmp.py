import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import imagegen, fpng_b200
from oracle.pyoracle import Oracle
o = Oracle(); fpng_b200.fpng_init()
SHAPES = [(1, 1), (16, 2), (20, 1), (24, 1), (64, 3), (128, 4), (256, 3), (512, 64), (528, 3), (1024, 5), (1040, 2), (1920, 3), (4096, 2), (8192, 1)]
for c in (4, 3):
  for kind in ('g0','g1'):
    for i,(w,h) in enumerate(SHAPES):
        img = imagegen.make(kind, w, h, c, 100 + i)
        for flags in (0,1,2):
            try:
                ok, png = fpng_b200.fpng_encode_image_to_memory(img, w, h, c, flags)
            except Exception as e:
                print("EXC", kind, w, h, c, flags, e); sys.exit(1)
            exp = o.encode(img, w, h, c, flags)
            if png != exp:
                n = min(len(png), len(exp)); d = next((j for j in range(n) if png[j] != exp[j]), n)
                print("MISMATCH", kind, w, h, c, flags, len(png), len(exp), "first diff", d)
print("done")
