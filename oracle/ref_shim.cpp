// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" wrappers around the UNMODIFIED reference sources that live in
// /root/reference/src (compiled from where they lie by oracle/Makefile; nothing is
// copied into this repository).  The resulting oracle/_ref/libfpng_ref.so is used by
// tests/ and by bench.py's cpu_baseline / --impl reference legs as
//   * the reference encoder/decoder (fpng.cpp, namespace fpng),
//   * independent PNG verifiers: lodepng (checks IDAT CRC + Adler-32) and stb_image.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "fpng.h"
#include "lodepng.h"

#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_PNG
#define STBI_NO_STDIO
#include "stb_image.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_init(void) { fpng::fpng_init(); }
REF_API int ref_cpu_supports_sse41(void) { return fpng::fpng_cpu_supports_sse41() ? 1 : 0; }
REF_API uint32_t ref_crc32(const void* p, size_t n, uint32_t prev) { return fpng::fpng_crc32(p, n, prev); }
REF_API uint32_t ref_adler32(const void* p, size_t n, uint32_t prev) { return fpng::fpng_adler32(p, n, prev); }

// Returns encoded size (0 on failure). If out==NULL or cap too small only the size is returned.
REF_API size_t ref_encode(const void* img, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags,
                          void* out, size_t cap)
{
    std::vector<uint8_t> buf;
    if (!fpng::fpng_encode_image_to_memory(img, w, h, chans, buf, flags)) return 0;
    if (out && cap >= buf.size()) memcpy(out, buf.data(), buf.size());
    return buf.size();
}

// Encode-only timing helper: encodes `reps` times and discards output (keeps the vector
// allocation inside the loop, as fpng_test does).  Returns last size.
REF_API size_t ref_encode_discard(const void* img, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags, int reps)
{
    size_t s = 0;
    for (int i = 0; i < reps; i++) {
        std::vector<uint8_t> buf;
        if (!fpng::fpng_encode_image_to_memory(img, w, h, chans, buf, flags)) return 0;
        s = buf.size();
    }
    return s;
}

// Decode-only timing helper (decodes `reps` times, keeps the vector allocation inside the loop like fpng_test).
REF_API int ref_decode_discard(const void* file, uint32_t size, uint32_t desired, int reps)
{
    int st = 0;
    for (int i = 0; i < reps; i++) {
        std::vector<uint8_t> buf; uint32_t w, h, c;
        st = fpng::fpng_decode_memory(file, size, buf, w, h, c, desired);
        if (st) return st;
    }
    return st;
}

REF_API int ref_get_info(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans)
{
    return fpng::fpng_get_info(file, size, *w, *h, *chans);
}

REF_API int ref_decode(const void* file, uint32_t size, void* out, size_t cap,
                       uint32_t* w, uint32_t* h, uint32_t* chans, uint32_t desired)
{
    std::vector<uint8_t> buf;
    int st = fpng::fpng_decode_memory(file, size, buf, *w, *h, *chans, desired);
    if (st == 0 && out && cap >= buf.size()) memcpy(out, buf.data(), buf.size());
    return st;
}

// lodepng (verifies CRC + Adler by default). want_chans 3 -> LCT_RGB, 4 -> LCT_RGBA. Returns lodepng error (0 ok).
REF_API unsigned ref_lodepng_decode(const void* file, size_t size, void* out, size_t cap,
                                    uint32_t* w, uint32_t* h, uint32_t want_chans)
{
    unsigned char* px = nullptr; unsigned ww = 0, hh = 0;
    unsigned err = lodepng_decode_memory(&px, &ww, &hh, (const unsigned char*)file, size,
                                         want_chans == 3 ? LCT_RGB : LCT_RGBA, 8);
    if (!err) {
        *w = ww; *h = hh;
        size_t n = (size_t)ww * hh * want_chans;
        if (out && cap >= n) memcpy(out, px, n);
    }
    free(px);
    return err;
}

// stb_image. Returns channels in file (0 on failure).
REF_API int ref_stb_decode(const void* file, int size, void* out, size_t cap,
                           uint32_t* w, uint32_t* h, int want_chans)
{
    int x = 0, y = 0, comp = 0;
    unsigned char* px = stbi_load_from_memory((const stbi_uc*)file, size, &x, &y, &comp, want_chans);
    if (!px) return 0;
    *w = (uint32_t)x; *h = (uint32_t)y;
    size_t n = (size_t)x * y * want_chans;
    if (out && cap >= n) memcpy(out, px, n);
    stbi_image_free(px);
    return comp;
}
