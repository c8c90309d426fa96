#!/usr/bin/env python3
"""tools/train_tables.py -- the counterpart of the reference's table-training driver (`fpng_test -t @filelist.txt`,
src/fpng_test.cpp:766-963) on the B200 path.

For every PNG of the listing: decode to RGBA with a general PNG decoder, treat the image as 32bpp iff some alpha < 255 (otherwise
24bpp; fpng_test.cpp:808-821), add its 16-bit scaled literal/length histogram (what the reference accumulates in g_huff_counts,
src/fpng.cpp:751-755) to the opaque or the alpha totals -- the histogram runs in the 2-pass histogram kernel on the GPU
(fpngb_train_accumulate_device) -- and re-verify the image through a 2-pass encode + independent decode like the reference does
(fpng_test.cpp:844-862).  Then fpngb_create_dynamic_block_prefix (src/fpng.cpp:909-988) turns each total into the pre-serialised
block header + code table, printed as the C source text the reference prints (so it can be pasted over g_dyn_huff_3/4 in
src/fpng.cpp:532-562), and optionally written as JSON for fpngb_set_static_table / fpng_b200.set_static_table.

    python tools/train_tables.py @filelist.txt [--json tables.json] [--check]

--check installs each trained table, encodes every file 1-pass with it and with the shipped table, verifies the pixels and prints
the size totals.  Needs a CUDA device (no CPU fallback); the listing format is the reference's (one file name per line).
"""
from __future__ import annotations

import argparse
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_listing(arg: str):
    """`@file`: one file name per line, trailing blanks stripped, empty lines skipped (fpng_test.cpp:279-350)."""
    if not arg.startswith("@"):
        raise SystemExit("Must specify list of files to read using @filelist.txt")
    names = []
    try:
        f = open(arg[1:], "r")
    except OSError:
        print(f'Failed opening listing file: "{arg[1:]}"', file=sys.stderr)          # fpng_test.cpp:282-290
        raise SystemExit(1)
    with f:
        for line in f:
            line = line.rstrip(" \n\r")
            if line:
                names.append(line)
    print(f'Successfully read {len(names)} filenames(s) from listing file "{arg[1:]}"')
    return names


def decode_rgba(data: bytes):
    """General PNG decode to 8-bit RGBA (the reference uses lodepng with LCT_RGBA, 8).  Returns [h, w, 4] uint8 or None."""
    from PIL import Image
    try:
        im = Image.open(io.BytesIO(data))
        im.load()
        return np.ascontiguousarray(np.asarray(im.convert("RGBA"), dtype=np.uint8))
    except Exception:
        return None


class GpuBackend:
    """The product path: histogram kernel + prefix builder of libfpng_b200.so."""

    def __init__(self, device: int = -1):
        import fpng_b200
        fpng_b200.fpng_init(device)
        self.f = fpng_b200

    def accumulate(self, img: np.ndarray, counts: np.ndarray) -> None:          # img [h, w, chans] uint8; counts uint64[288], in place
        import torch
        self.f.train_accumulate_device(torch.from_numpy(img[None]).cuda().contiguous(), counts)

    def create_prefix(self, counts: np.ndarray, chans: int):
        return self.f.create_dynamic_block_prefix(counts, chans)

    def sanity(self, img: np.ndarray) -> bool:
        """2-pass encode, then an independent decode must return the pixels (fpng_test.cpp:844-862)."""
        h, w, c = img.shape
        ok, png = self.f.fpng_encode_image_to_memory(img, w, h, c, 1)
        if not ok:
            return False
        back = decode_rgba(png)
        return back is not None and np.array_equal(back[:, :, :c], img) and (c == 4 or bool((back[:, :, 3] == 255).all()))

    def sizes_with_table(self, images, chans: int, table) -> int:
        total = 0
        try:
            if table is not None:
                self.f.set_static_table(chans, table[0], table[1], table[2])
            for img in images:
                h, w, c = img.shape
                ok, png = self.f.fpng_encode_image_to_memory(img, w, h, c, 0)
                st, px, *_ = self.f.fpng_decode_memory(png, c)
                if not ok or st != 0 or not np.array_equal(px, img.reshape(-1)):
                    raise SystemExit("1-pass encode/decode verification failed under the trained table")
                total += len(png)
        finally:
            if table is not None:
                self.f.set_static_table(chans)
        return total


def c_source(chans: int, prefix: bytes, bit_buf: int, bit_buf_size: int, codes, sizes) -> str:
    """The text `fpng_test -t` prints for one table (fpng_test.cpp:905-925 / 941-961), character for character."""
    out = ["\n", f"static const uint8_t g_dyn_huff_{chans}[] = {{\n"]
    n = len(prefix)
    for i, b in enumerate(prefix):
        out.append(f"{b}{',' if i != n - 1 else ' '} ")
        if (i & 31) == 31:
            out.append("\n")
    out.append("};\n")
    out.append(f"const uint32_t DYN_HUFF_{chans}_BITBUF = {bit_buf & 0xFFFFFFFF}, DYN_HUFF_{chans}_BITBUF_SIZE = {bit_buf_size};\n")
    out.append(f"static const struct {{ uint8_t m_code_size; uint16_t m_code; }} g_dyn_huff_{chans}_codes[288] = {{\n")
    for i in range(288):
        out.append(f"{{{int(sizes[i])},{int(codes[i])}}}{',' if i != 287 else ' '}")
        if (i & 31) == 31:
            out.append("\n")
    out.append("};\n")
    return "".join(out)


def train(files, backend, out=sys.stdout, check: bool = False):
    """Returns {3: table or None, 4: table or None}; table = (prefix, bit_buf, bit_buf_size, codes, sizes)."""
    freq = {3: np.zeros(288, np.uint64), 4: np.zeros(288, np.uint64)}
    nfiles = {3: 0, 4: 0}
    kept = {3: [], 4: []}
    failed = 0
    for name in files:
        print(f'Processing file "{name}"', file=out)
        try:
            with open(name, "rb") as f:
                data = f.read()
        except OSError:
            print(f'Failed reading source file data "{name}"', file=sys.stderr)
            raise SystemExit(1)
        rgba = decode_rgba(data)
        if rgba is None:
            print(f'WARNING: Failed unpacking source file "{name}"! Skipping.', file=sys.stderr)
            failed += 1
            continue
        h, w = rgba.shape[:2]
        has_alpha = bool((rgba[:, :, 3] < 255).any())
        chans = 4 if has_alpha else 3
        total = w * h
        print(f"Dimensions: {w}x{h}, Has Alpha: {int(has_alpha)}, Total Pixels: {total}, bytes: {total * chans} "
              f"({np.float32(total * chans) / np.float32(1024.0 * 1024.0):f} MB)", file=out)
        img = rgba if has_alpha else np.ascontiguousarray(rgba[:, :, :3])
        backend.accumulate(img, freq[chans])
        if not backend.sanity(img):
            print("FPNG decode verification failed!", file=sys.stderr)
            raise SystemExit(1)
        nfiles[chans] += 1
        if check:
            kept[chans].append(img)
    print(f"Total alpha files: {nfiles[4]}", file=out)
    print(f"Total opaque files: {nfiles[3]}", file=out)
    print(f"Total failed loading: {failed}", file=out)
    if not nfiles[3] and not nfiles[4]:
        print("No files were loaded!", file=sys.stderr)
        raise SystemExit(1)
    tables = {3: None, 4: None}
    for chans in (3, 4):                                  # the reference prints the opaque table first
        if not nfiles[chans]:
            continue
        t = backend.create_prefix(freq[chans], chans)
        tables[chans] = t
        out.write(c_source(chans, t[0], t[1], t[2], t[3], t[4]))
    if check:
        for chans in (3, 4):
            if tables[chans] is None:
                continue
            a = backend.sizes_with_table(kept[chans], chans, None)
            b = backend.sizes_with_table(kept[chans], chans, tables[chans])
            print(f"// {chans * 8}bpp: {len(kept[chans])} files, 1-pass bytes with the shipped table {a}, with the trained table {b} ({100.0 * b / a:.2f} %)", file=out)
    return tables


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("listing", help="@filelist.txt (one PNG per line), like fpng_test -t")
    ap.add_argument("--json", help="also write the tables as JSON (prefix bytes, bit_buf, bit_buf_size, code sizes, codes)")
    ap.add_argument("--check", action="store_true", help="install each trained table and compare 1-pass sizes on the training files")
    ap.add_argument("--device", type=int, default=-1)
    args = ap.parse_args(argv)
    files = load_listing(args.listing)
    tables = train(files, GpuBackend(args.device), check=args.check)
    if args.json:
        doc = {str(c): {"prefix": list(t[0]), "bit_buf": int(t[1]), "bit_buf_size": int(t[2]),
                        "codes": [int(v) for v in t[3]], "sizes": [int(v) for v in t[4]]} for c, t in tables.items() if t is not None}
        with open(args.json, "w") as f:
            json.dump(doc, f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
