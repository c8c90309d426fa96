// fpng_b200/csrc/decode_api.cu -- container walk (host) and decode entry points of the C ABI.
#include "../../include/fpng_b200.h"
#include "kernels.cuh"
#include <string.h>

namespace fpngb {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// PNG signature / IHDR / chunk walk with the reference's acceptance rules and return codes
// (fpng.cpp:2930-3077).  Only chunk headers and the CRCs of the small non-IDAT chunks are touched: this is the
// part of the container that stays on the host.
int container_info(const uint8_t* f, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans, uint32_t* idat_ofs, uint32_t* idat_len)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    *w = *h = *chans = 0; *idat_ofs = *idat_len = 0;
    if (size < 8 + 25 + 8 + 1 + 4 + 12) return FPNGB_DECODE_FAILED_NOT_PNG;
    if (memcmp(f, sig, 8) != 0) return FPNGB_DECODE_FAILED_NOT_PNG;
    if (be32(f + 8) != 13) return FPNGB_DECODE_FAILED_NOT_PNG;
    if (host_crc32(f + 12, 17, 0) != be32(f + 29)) return FPNGB_DECODE_FAILED_HEADER_CRC32;
    *w = be32(f + 16); *h = be32(f + 20);
    if (!*w || !*h || *w > (1u << 24) || *h > (1u << 24)) return FPNGB_DECODE_FAILED_INVALID_DIMENSIONS;
    if ((uint64_t)*w * *h > (1u << 30)) return FPNGB_DECODE_FAILED_INVALID_DIMENSIONS;
    if (f[26] || f[27] || f[28] || f[24] != 8) return FPNGB_DECODE_NOT_FPNG;
    if (f[25] == 2) *chans = 3; else if (f[25] == 6) *chans = 4;
    if (!*chans) return FPNGB_DECODE_NOT_FPNG;
    bool have_fdec = false;
    size_t ofs = 33;
    for (;;) {
        if (ofs >= size || size - ofs < 12) return FPNGB_DECODE_FAILED_CHUNK_PARSING;
        const uint32_t len = be32(f + ofs);
        if ((uint64_t)ofs + 8 + len + 4 > size) return FPNGB_DECODE_FAILED_CHUNK_PARSING;
        const uint8_t* ty = f + ofs + 4;
        for (int i = 0; i < 4; i++) {
            const bool up = ty[i] >= 65 && ty[i] <= 90, lo = ty[i] >= 97 && ty[i] <= 122;
            if (!up && !lo) return FPNGB_DECODE_FAILED_CHUNK_PARSING;
        }
        const bool is_idat = memcmp(ty, "IDAT", 4) == 0;
        if (!is_idat && host_crc32(ty, 4 + (size_t)len, 0) != be32(f + ofs + 8 + len)) return FPNGB_DECODE_FAILED_HEADER_CRC32;
        const uint8_t* d = f + ofs + 8;
        if (memcmp(ty, "IEND", 4) == 0) break;
        if (is_idat) {
            if (*idat_ofs || !have_fdec) return FPNGB_DECODE_NOT_FPNG;
            *idat_ofs = (uint32_t)ofs; *idat_len = len;
            if (len < 7) return FPNGB_DECODE_FAILED_INVALID_IDAT;
        } else if (memcmp(ty, "fdEC", 4) == 0) {
            if (have_fdec || len != 5) return FPNGB_DECODE_NOT_FPNG;
            if (d[0] != 82 || d[1] != 36 || d[2] != 147 || d[3] != 227 || d[4] != 0) return FPNGB_DECODE_NOT_FPNG;
            have_fdec = true;
        } else if (!(ty[0] & 32)) return FPNGB_DECODE_NOT_FPNG;
        ofs += 8 + (size_t)len + 4;
    }
    if (!have_fdec || !*idat_ofs) return FPNGB_DECODE_NOT_FPNG;
    return FPNGB_DECODE_SUCCESS;
}

}  // namespace fpngb

using namespace fpngb;

extern "C" {

int fpngb_get_info(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans)
{
    uint32_t a, b, ww = 0, hh = 0, cc = 0;
    if (!file) { if (w) *w = 0; if (h) *h = 0; if (chans) *chans = 0; return FPNGB_DECODE_INVALID_ARG; }
    const int st = container_info((const uint8_t*)file, size, &ww, &hh, &cc, &a, &b);
    if (w) *w = ww; if (h) *h = hh; if (chans) *chans = cc;
    return st;
}

}  // extern "C"
