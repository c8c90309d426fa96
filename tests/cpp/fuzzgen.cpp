// tests/cpp/fuzzgen.cpp -- TEST INFRASTRUCTURE.  Input generators of the reference's two encoder fuzzers, restated so
// that the SAME SEEDS give the SAME inputs as `fpng_test -e` / `fpng_test -E` built with this toolchain's libstdc++:
//   * fuzzgen_mutate():  one trial of fuzz_test_encoder (src/fpng_test.cpp:381-510): std::mt19937 seeded with the
//                        trial number drives the family choice and the run structure, glibc rand() seeded with the
//                        trial number drives the "full random" and "bit flip" families;
//   * fuzzgen_dims_*():  fuzz_test_encoder2 (src/fpng_test.cpp:617-645): ONE default-seeded (5489) std::mt19937 for the
//                        whole session: width, height in [1, 8194], a coin for 3/4 channels, then one 32-bit draw per pixel.
// Only the input generation is restated here; encoding, decoding and comparing happen in the python tests.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <random>

#define FZ_API extern "C" __attribute__((visibility("default")))

namespace {
struct Rng {
    std::mt19937 mt;
    int between(int lo, int hi) { std::uniform_int_distribution<int> d(lo, hi); return d(mt); }
    uint32_t u32() { return (uint32_t)between(INT32_MIN, INT32_MAX); }
    float unit(float lo, float hi) { std::uniform_real_distribution<float> d(lo, hi); return d(mt); }
};
}

// Applies trial `trial`'s mutation to buf (n bytes of chans-channel pixels, already holding the source image).
// Returns the family: 0 colour-fill runs, 1 colour-xor runs, 2 byte-fill runs, 3 byte-xor runs, 4 full random, 5 bit flips.
FZ_API int fuzzgen_mutate(uint32_t trial, uint8_t* buf, uint32_t n, uint32_t chans)
{
    Rng r; r.mt.seed(trial);
    srand(trial);
    const double fract = r.unit(0.000001f, .1f);
    const uint32_t thresh = (uint32_t)((double)RAND_MAX * fract);
    auto pixel_runs = [&](bool xor_mode) {
        for (uint32_t o = 0; o < n; ) {
            const uint32_t left_px = (n - o) / chans;
            const uint32_t len = (uint32_t)r.between(1, (int)std::min<uint32_t>(left_px, 32));
            uint8_t c[4];
            for (int k = 0; k < 4; k++) c[k] = (uint8_t)r.between(0, 0xFF);
            if (!xor_mode) {
                for (uint32_t i = 0; i < len; i++, o += chans) memcpy(buf + o, c, chans);
            } else if (r.unit(0.0f, 1.0f) > .8f) {
                for (uint32_t i = 0; i < len; i++, o += chans) for (uint32_t j = 0; j < chans; j++) buf[o + j] ^= c[j];
            } else o += len * chans;
        }
    };
    if (r.unit(0.0f, 1.0f) < .05f) { pixel_runs(false); return 0; }
    if (r.unit(0.0f, 1.0f) < .05f) { pixel_runs(true); return 1; }
    if (r.unit(0.0f, 1.0f) < .05f) {
        for (uint32_t o = 0; o < n; ) {
            const uint32_t len = (uint32_t)r.between(1, (int)std::min<uint32_t>(n - o, 258));
            const int v = r.between(0, 0xFF);
            memset(buf + o, v, len);
            o += len;
        }
        return 2;
    }
    if (r.unit(0.0f, 1.0f) < .15f) {
        for (uint32_t o = 0; o < n; ) {
            const uint32_t len = (uint32_t)r.between(1, (int)std::min<uint32_t>(n - o, 32));
            if (r.unit(0.0f, 1.0f) > .1f) {
                const uint32_t v = (uint32_t)r.between(0, 0xFF);
                for (uint32_t i = 0; i < len; i++) buf[o + i] ^= (uint8_t)v;
            }
            o += len;
        }
        return 3;
    }
    if (r.unit(0.0f, 1.0f) < .005f) {
        for (uint32_t i = 0; i < n; i++) buf[i] = (uint8_t)rand();
        return 4;
    }
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = 0; j < 8; j++)
            if ((uint32_t)rand() <= thresh) buf[i] ^= (uint8_t)(1u << j);
    return 5;
}

// fuzz_test_encoder2: a session object walking the trials in order.
struct DimSession { Rng r; };
FZ_API void* fuzzgen_dims_open(void) { return new DimSession(); }            // std::mt19937 default seed 5489
FZ_API void fuzzgen_dims_close(void* s) { delete (DimSession*)s; }
// draws the next trial's header; the caller then MUST call fuzzgen_dims_fill (or _skip) before the next header
FZ_API void fuzzgen_dims_next(void* s, uint32_t* w, uint32_t* h, uint32_t* chans)
{
    DimSession* d = (DimSession*)s;
    *w = (uint32_t)d->r.between(1, 8193 + 1);
    *h = (uint32_t)d->r.between(1, 8193 + 1);
    *chans = d->r.between(0, 1) == 1 ? 4 : 3;
}
FZ_API void fuzzgen_dims_fill(void* s, uint8_t* dst, uint64_t pixels, uint32_t chans)
{
    DimSession* d = (DimSession*)s;
    for (uint64_t i = 0; i < pixels; i++) {
        const uint32_t p = d->r.u32();
        *dst++ = (uint8_t)p; *dst++ = (uint8_t)(p >> 8); *dst++ = (uint8_t)(p >> 16);
        if (chans == 4) *dst++ = (uint8_t)(p >> 24);
    }
}
FZ_API void fuzzgen_dims_skip(void* s, uint64_t pixels)
{
    DimSession* d = (DimSession*)s;
    for (uint64_t i = 0; i < pixels; i++) (void)d->r.u32();
}
