"""Writes tests/golden/static_tables.json: the 1-pass Huffman tables exactly as the reference stores them
(/root/reference/src/fpng.cpp: g_dyn_huff_3 / g_dyn_huff_4 header bytes with their bit-buffer tails, and the
g_dyn_huff_3_codes / g_dyn_huff_4_codes {size, code} pairs), read from the reference SOURCE TEXT -- data of the file format,
not code.  The product derives its code books at init from the header bytes alone (csrc/static_tables.h, host_api.cu
build_static_book); tests/test_container_cpu.py pins them to this fixture, independently of the oracle.
Run in the build container (needs /root/reference): python tests/golden/make_static_tables.py"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
text = open("/root/reference/src/fpng.cpp").read()
out = {}
for chans in (3, 4):
    m = re.search(r"g_dyn_huff_%d\[\] = \{(.*?)\};" % chans, text, re.S)
    hdr = [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]
    bb = int(re.search(r"DYN_HUFF_%d_BITBUF = (0x[0-9a-fA-F]+|\d+)" % chans, text).group(1), 0)
    bs = int(re.search(r"DYN_HUFF_%d_BITBUF_SIZE = (\d+)" % chans, text).group(1))
    m = re.search(r"g_dyn_huff_%d_codes\[288\] = \{(.*?)\};" % chans, text, re.S)
    pairs = [(int(s), int(c)) for s, c in re.findall(r"\{(\d+),(\d+)\}", m.group(1))]
    assert len(pairs) == 288
    out[str(chans)] = {"header_bytes": hdr, "bit_buf": bb, "bit_buf_size": bs, "sizes": [p[0] for p in pairs], "codes": [p[1] for p in pairs]}
json.dump(out, open(os.path.join(HERE, "static_tables.json"), "w"))
print({k: (len(v["header_bytes"]), v["bit_buf"], v["bit_buf_size"]) for k, v in out.items()})
