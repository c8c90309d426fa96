// tests/cpp/dropin_test.cpp -- a C++ caller written against the reference's public header (namespace fpng), linked with
// libfpng_b200.so instead of fpng.cpp.  Mirrors the reference harness's checks (src/fpng_test.cpp:1237-1327: encode,
// decode with fpng, memcmp; 4->3 and 3->4 conversion) and its uniform-random fuzz shape family (617-682).
#include "fpng.h"

#include <stdio.h>
#include <string.h>
#include <vector>

static uint32_t g_state = 12345;
static uint32_t lcg() { g_state = g_state * 1664525u + 1013904223u; return g_state >> 8; }

static int check_image(uint32_t w, uint32_t h, uint32_t chans, int kind, uint32_t flags)
{
    std::vector<uint8_t> img((size_t)w * h * chans);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++)
            for (uint32_t c = 0; c < chans; c++) {
                uint32_t v = kind == 0 ? lcg() : ((x * 3 + y * 2 + c * 40) / 4 + (kind == 1 ? (lcg() % 5) : 0));
                img[((size_t)y * w + x) * chans + c] = (uint8_t)v;
            }
    std::vector<uint8_t> png;
    if (!fpng::fpng_encode_image_to_memory(img.data(), w, h, chans, png, flags)) { printf("encode failed %ux%ux%u\n", w, h, chans); return 1; }
    uint32_t W = 0, H = 0, C = 0;
    if (fpng::fpng_get_info(png.data(), (uint32_t)png.size(), W, H, C) != fpng::FPNG_DECODE_SUCCESS || W != w || H != h || C != chans) { printf("get_info mismatch\n"); return 1; }
    for (uint32_t desired = 3; desired <= 4; desired++) {
        std::vector<uint8_t> out;
        int st = fpng::fpng_decode_memory(png.data(), (uint32_t)png.size(), out, W, H, C, desired);
        if (st != fpng::FPNG_DECODE_SUCCESS || out.size() != (size_t)w * h * desired) { printf("decode failed st=%d\n", st); return 1; }
        for (size_t p = 0; p < (size_t)w * h; p++) {
            for (uint32_t c = 0; c < 3; c++) if (out[p * desired + c] != img[p * chans + c]) { printf("pixel mismatch\n"); return 1; }
            if (desired == 4 && out[p * 4 + 3] != (chans == 4 ? img[p * 4 + 3] : 0xFF)) { printf("alpha mismatch\n"); return 1; }
        }
    }
    return 0;
}

int main()
{
    fpng::fpng_init();
    if (fpng::fpng_crc32("123456789", 9) != 0xCBF43926u) { printf("crc32 KAT failed\n"); return 1; }
    if (fpng::fpng_adler32("Wikipedia", 9) != 0x11E60398u) { printf("adler32 KAT failed\n"); return 1; }
    std::vector<uint8_t> tmp;
    uint8_t px[12] = {0};
    if (fpng::fpng_encode_image_to_memory(px, 0, 1, 3, tmp)) { printf("w=0 accepted\n"); return 1; }
    if (fpng::fpng_encode_image_to_memory(px, 2, 2, 5, tmp)) { printf("chans=5 accepted\n"); return 1; }
    uint32_t W, H, C;
    if (fpng::fpng_decode_memory(px, 12, tmp, W, H, C, 7) != fpng::FPNG_DECODE_INVALID_ARG) { printf("desired=7 accepted\n"); return 1; }
    if (fpng::fpng_decode_memory(px, 12, tmp, W, H, C, 3) != fpng::FPNG_DECODE_FAILED_NOT_PNG) { printf("garbage accepted\n"); return 1; }
    int fails = 0, n = 0;
    const uint32_t shapes[][2] = {{1, 1}, {3, 2}, {17, 5}, {64, 64}, {129, 33}, {512, 200}, {1000, 7}, {1920, 16}};
    for (auto& s : shapes)
        for (uint32_t chans = 3; chans <= 4; chans++)
            for (int kind = 0; kind < 3; kind++)
                for (uint32_t flags = 0; flags < 3; flags++) { fails += check_image(s[0], s[1], chans, kind, flags); n++; }
    // random dimensions like fuzz_test_encoder2
    for (int t = 0; t < 20; t++) { fails += check_image(1 + lcg() % 700, 1 + lcg() % 40, (lcg() & 1) ? 4 : 3, 0, 0); n++; }
    if (!fpng::fpng_encode_image_to_file("/tmp/fpng_b200_dropin.png", px, 2, 2, 3)) { printf("to_file failed\n"); return 1; }
    std::vector<uint8_t> back;
    if (fpng::fpng_decode_file("/tmp/fpng_b200_dropin.png", back, W, H, C, 3) != fpng::FPNG_DECODE_SUCCESS || back.size() != 12) { printf("decode_file failed\n"); return 1; }
    printf("dropin_test: %d cases, %d failures\n", n, fails);
    return fails ? 1 : 0;
}
