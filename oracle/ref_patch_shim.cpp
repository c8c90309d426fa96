// oracle/ref_patch_shim.cpp -- TEST INFRASTRUCTURE ONLY.  The reference encoder with PATCHABLE 1-pass tables.
//
// The reference's two trained 1-pass tables (src/fpng.cpp:532-562: g_dyn_huff_{3,4}, DYN_HUFF_{3,4}_BITBUF[_SIZE],
// g_dyn_huff_{3,4}_codes) are `static const`.  This translation unit compiles the reference source WHERE IT LIES
// (#include of /root/reference/src/fpng.cpp, nothing copied) with the keyword `const` defined away, which turns those
// tables into ordinary mutable globals; refp_set_table() then overwrites them in memory with a table produced by the
// reference's own trainer (create_dynamic_block_prefix, src/fpng.cpp:909-988).  Every line of the encoder is the
// reference's; only the table CONTENTS change.  Purpose: byte-parity tests of fpngb_set_static_table(), in particular
// of the RGBA "1-pixel match vs 4 literals" rule (src/fpng.cpp:1520-1528), which is dead code under the shipped table.
// Limitation: the prefix byte length is a compile-time sizeof in the reference (62 RGB / 61 RGBA), so only trained
// prefixes of exactly that length can be installed (tests search for one).
// Built into oracle/_ref/libfpng_ref_patch.so by oracle/Makefile (needs -fpermissive because const is gone).
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <assert.h>
#include <vector>
#include <algorithm>
#include <math.h>
#include <smmintrin.h>
#include <wmmintrin.h>
#include <emmintrin.h>
#include <xmmintrin.h>
#if defined(__GNUC__)
#include <cpuid.h>
#endif
#define const
#include "fpng.cpp"
#undef const

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct Saved { uint8_t hdr[64]; uint32_t bitbuf, bitbuf_size; uint8_t sizes[288]; uint16_t codes[288]; bool valid; };
Saved g_saved[2];
void save_once()
{
    if (g_saved[0].valid) return;
    memcpy(g_saved[0].hdr, fpng::g_dyn_huff_3, sizeof(fpng::g_dyn_huff_3));
    g_saved[0].bitbuf = fpng::DYN_HUFF_3_BITBUF; g_saved[0].bitbuf_size = fpng::DYN_HUFF_3_BITBUF_SIZE;
    memcpy(g_saved[1].hdr, fpng::g_dyn_huff_4, sizeof(fpng::g_dyn_huff_4));
    g_saved[1].bitbuf = fpng::DYN_HUFF_4_BITBUF; g_saved[1].bitbuf_size = fpng::DYN_HUFF_4_BITBUF_SIZE;
    for (int i = 0; i < 288; i++) {
        g_saved[0].sizes[i] = fpng::g_dyn_huff_3_codes[i].m_code_size; g_saved[0].codes[i] = fpng::g_dyn_huff_3_codes[i].m_code;
        g_saved[1].sizes[i] = fpng::g_dyn_huff_4_codes[i].m_code_size; g_saved[1].codes[i] = fpng::g_dyn_huff_4_codes[i].m_code;
    }
    g_saved[0].valid = g_saved[1].valid = true;
}
}

REF_API void refp_init(void) { fpng::fpng_init(); save_once(); }
REF_API uint32_t refp_table_len(uint32_t chans) { return chans == 3 ? (uint32_t)sizeof(fpng::g_dyn_huff_3) : (uint32_t)sizeof(fpng::g_dyn_huff_4); }

// installs a trained table; returns 1 on success, 0 if the prefix length does not equal the reference's compile-time size
REF_API int refp_set_table(uint32_t chans, uint8_t* prefix, size_t len, uint32_t bit_buf, uint32_t bit_buf_size, uint32_t* codes288, uint8_t* sizes288)
{
    save_once();
    if (len != refp_table_len(chans)) return 0;
    if (chans == 3) {
        memcpy(fpng::g_dyn_huff_3, prefix, len); fpng::DYN_HUFF_3_BITBUF = bit_buf; fpng::DYN_HUFF_3_BITBUF_SIZE = bit_buf_size;
        for (int i = 0; i < 288; i++) { fpng::g_dyn_huff_3_codes[i].m_code_size = sizes288[i]; fpng::g_dyn_huff_3_codes[i].m_code = (uint16_t)codes288[i]; }
    } else {
        memcpy(fpng::g_dyn_huff_4, prefix, len); fpng::DYN_HUFF_4_BITBUF = bit_buf; fpng::DYN_HUFF_4_BITBUF_SIZE = bit_buf_size;
        for (int i = 0; i < 288; i++) { fpng::g_dyn_huff_4_codes[i].m_code_size = sizes288[i]; fpng::g_dyn_huff_4_codes[i].m_code = (uint16_t)codes288[i]; }
    }
    return 1;
}

REF_API void refp_reset_table(uint32_t chans)
{
    save_once();
    Saved& s = g_saved[chans == 3 ? 0 : 1];
    if (chans == 3) {
        memcpy(fpng::g_dyn_huff_3, s.hdr, sizeof(fpng::g_dyn_huff_3)); fpng::DYN_HUFF_3_BITBUF = s.bitbuf; fpng::DYN_HUFF_3_BITBUF_SIZE = s.bitbuf_size;
        for (int i = 0; i < 288; i++) { fpng::g_dyn_huff_3_codes[i].m_code_size = s.sizes[i]; fpng::g_dyn_huff_3_codes[i].m_code = s.codes[i]; }
    } else {
        memcpy(fpng::g_dyn_huff_4, s.hdr, sizeof(fpng::g_dyn_huff_4)); fpng::DYN_HUFF_4_BITBUF = s.bitbuf; fpng::DYN_HUFF_4_BITBUF_SIZE = s.bitbuf_size;
        for (int i = 0; i < 288; i++) { fpng::g_dyn_huff_4_codes[i].m_code_size = s.sizes[i]; fpng::g_dyn_huff_4_codes[i].m_code = s.codes[i]; }
    }
}

REF_API size_t refp_encode(void* img, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags, uint8_t* out, size_t cap)
{
    std::vector<uint8_t> buf;
    if (!fpng::fpng_encode_image_to_memory(img, w, h, chans, buf, flags)) return 0;
    if (buf.size() > cap) return buf.size();
    memcpy(out, buf.data(), buf.size());
    return buf.size();
}

REF_API int refp_decode(void* data, uint32_t size, uint8_t* out, size_t cap, uint32_t* w, uint32_t* h, uint32_t* c, uint32_t desired)
{
    std::vector<uint8_t> px;
    int st = fpng::fpng_decode_memory(data, size, px, *w, *h, *c, desired);
    if (st == 0 && px.size() <= cap) memcpy(out, px.data(), px.size());
    return st;
}
