// tools/fpng_b200_test.cpp -- command-line harness for the B200 fpng hot path, the counterpart of the reference's
// fpng_test (src/fpng_test.cpp:975-1639) restricted to the fpng columns: it loads a PNG, encodes it through the drop-in
// C++ API (include/fpng.h -> libfpng_b200.so -> CUDA kernels), writes fpng.png, decodes it again, verifies the pixels,
// checks the 4->3 / 3->4 channel conversions, verifies the file with an INDEPENDENT general PNG decoder (below: zlib
// inflate + the five PNG filters, written for this tool) and prints timings or a CSV row.
//
//   fpng_b200_test [options] file.png
//     -s  2-pass compression (FPNG_ENCODE_SLOWER)            (src/fpng_test.cpp:1015)
//     -u  stored Deflate blocks (FPNG_FORCE_UNCOMPRESSED)     (src/fpng_test.cpp:1011)
//     -a  swizzle green into alpha -> 32bpp test             (src/fpng_test.cpp:1147-1152)
//     -c  CSV row: file, w, h, chans, encode secs, MiB, decode secs, encode MP/s, decode MP/s   (the fpng columns of 1608-1633)
//     -f  decode the file with fpng and exit (decoder fuzzing entry, src/fpng_test.cpp:1092-1114)
//     -e  encoder fuzz: the six mutation families with the reference's seeds (src/fpng_test.cpp:381-615), N trials (-n)
//     -E  encoder fuzz 2: random dimensions / uniform random pixels, default-seeded mt19937 (src/fpng_test.cpp:617-682)
//     -n N  number of fuzz trials (default 1000 / 1000 like the reference; use fewer for a smoke run)
//     -o file  where to write the encoded file (default fpng.png)
// The comparison codecs of the reference harness (lodepng / stb_image_write / qoi encoders, wuffs, pvpng) are not part
// of the hot path and are not rebuilt here.
//
// General-PNG fallback: fpng_decode_memory() only reads fpng-written files; for anything else it returns
// FPNG_DECODE_NOT_FPNG so that the caller falls back to a general decoder (src/fpng.h:79-91).  This tool registers its
// zlib-based decoder through fpngb_set_fallback_decoder(), after which fpng::fpng_decode_memory/_file hand such files
// to it transparently -- the "fallback hook" of SURVEY.md section 8f.
#include "fpng.h"
#include "fpng_b200.h"

#include <zlib.h>

#include <chrono>
#include <random>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

typedef std::vector<uint8_t> bytes;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static bool read_file(const char* name, bytes& out)
{
    FILE* f = fopen(name, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    const bool ok = fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

// ---- a small general PNG decoder (8-bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced): the independent verifier
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static int paeth(int a, int b, int c) { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

static bool general_png_decode(const uint8_t* file, size_t size, uint32_t desired, bytes& out, uint32_t& w, uint32_t& h, uint32_t& chans_in_file, bool check_crc = true)
{
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (size < 8 + 25 || memcmp(file, sig, 8)) return false;
    size_t o = 8;
    bytes idat, plte;
    uint32_t depth = 0, ctype = 0, interlace = 0;
    bool have_ihdr = false, done = false;
    while (!done && o + 12 <= size) {
        const uint32_t len = be32(file + o);
        if ((uint64_t)o + 12 + len > size) return false;
        const uint8_t* type = file + o + 4; const uint8_t* data = file + o + 8;
        if (check_crc && (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != be32(data + len)) return false;
        if (!memcmp(type, "IHDR", 4)) { if (len != 13) return false; w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; have_ihdr = true; }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) done = true;
        o += 12 + (size_t)len;
    }
    if (!have_ihdr || !w || !h || depth != 8 || interlace) return false;
    const uint32_t comps = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!comps) return false;
    chans_in_file = (ctype == 4 || ctype == 6) ? 4 : 3;
    const size_t bpl = (size_t)w * comps;
    bytes raw((bpl + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return false;   // checks the Adler-32
    bytes img(bpl * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* src = raw.data() + (bpl + 1) * y; const uint8_t ft = src[0]; src++;
        uint8_t* dst = img.data() + bpl * y; const uint8_t* up = y ? dst - bpl : nullptr;
        for (size_t i = 0; i < bpl; i++) {
            const int a = i >= comps ? dst[i - comps] : 0, b = up ? up[i] : 0, c = (up && i >= comps) ? up[i - comps] : 0;
            int v = src[i];
            switch (ft) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break; default: return false; }
            dst[i] = (uint8_t)v;
        }
    }
    out.resize((size_t)w * h * desired);
    for (size_t p = 0; p < (size_t)w * h; p++) {
        uint8_t r, g, b, a = 255;
        const uint8_t* s = img.data() + p * comps;
        if (ctype == 0) { r = g = b = s[0]; }
        else if (ctype == 4) { r = g = b = s[0]; a = s[1]; }
        else if (ctype == 3) { if ((size_t)s[0] * 3 + 2 >= plte.size()) return false; r = plte[s[0] * 3]; g = plte[s[0] * 3 + 1]; b = plte[s[0] * 3 + 2]; }
        else { r = s[0]; g = s[1]; b = s[2]; if (ctype == 6) a = s[3]; }
        uint8_t* d = out.data() + p * desired;
        d[0] = r; d[1] = g; d[2] = b; if (desired == 4) d[3] = a;
    }
    return true;
}

// the fallback hook handed to the library (see fpngb_set_fallback_decoder in include/fpng_b200.h)
static int fallback_cb(const void* file, uint32_t size, uint32_t desired, void* /*user*/, void** pixels, uint32_t* w, uint32_t* h, uint32_t* chans)
{
    bytes px;
    if (!general_png_decode((const uint8_t*)file, size, desired, px, *w, *h, *chans)) return 1;
    *pixels = malloc(px.size());
    if (!*pixels) return 1;
    memcpy(*pixels, px.data(), px.size());
    return 0;
}

// ---- fuzz input generators: same seeds, same libstdc++ streams as the reference harness (see tests/cpp/fuzzgen.cpp)
extern "C" int fuzzgen_mutate(uint32_t trial, uint8_t* buf, uint32_t n, uint32_t chans);
extern "C" void* fuzzgen_dims_open(void);
extern "C" void fuzzgen_dims_close(void* s);
extern "C" void fuzzgen_dims_next(void* s, uint32_t* w, uint32_t* h, uint32_t* chans);
extern "C" void fuzzgen_dims_fill(void* s, uint8_t* dst, uint64_t pixels, uint32_t chans);

static bool roundtrip_ok(const bytes& img, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags, size_t* out_size)
{
    bytes png;
    if (!fpng::fpng_encode_image_to_memory(img.data(), w, h, chans, png, flags)) { fprintf(stderr, "fpng_encode_image_to_memory() failed!\n"); return false; }
    if (out_size) *out_size = png.size();
    bytes gen; uint32_t gw, gh, gc;
    if (!general_png_decode(png.data(), png.size(), chans, gen, gw, gh, gc) || gw != w || gh != h || gen != img) { fprintf(stderr, "general PNG decoder verification failure!\n"); return false; }
    bytes dec; uint32_t dw, dh, dc;
    const int st = fpng::fpng_decode_memory(png.data(), (uint32_t)png.size(), dec, dw, dh, dc, 4);
    if (st != fpng::FPNG_DECODE_SUCCESS || dw != w || dh != h || dc != chans) { fprintf(stderr, "fpng_decode_memory() failed with %d\n", st); return false; }
    for (size_t p = 0; p < (size_t)w * h; p++) {
        for (uint32_t j = 0; j < chans; j++) if (dec[p * 4 + j] != img[p * chans + j]) { fprintf(stderr, "fpng verification failure!\n"); return false; }
        if (chans == 3 && dec[p * 4 + 3] != 0xFF) { fprintf(stderr, "fpng verification failure (alpha)!\n"); return false; }
    }
    return true;
}

int main(int argc, char** argv)
{
    const char* filename = nullptr; const char* out_name = "fpng.png";
    bool csv = false, slower = false, uncompressed = false, fuzz_e = false, fuzz_E = false, decode_only = false, swizzle = false;
    uint32_t trials = 1000;
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
        if (a[0] == '-') {
            switch (a[1]) {
            case 'u': uncompressed = true; break; case 's': slower = true; break; case 'c': csv = true; break;
            case 'e': fuzz_e = true; break; case 'E': fuzz_E = true; break; case 'f': decode_only = true; break; case 'a': swizzle = true; break;
            case 'n': if (i + 1 < argc) trials = (uint32_t)atoi(argv[++i]); break;
            case 'o': if (i + 1 < argc) out_name = argv[++i]; break;
            default: fprintf(stderr, "Unrecognized option: %s\n", a); return EXIT_FAILURE;
            }
        } else filename = a;
    }
    if (!filename && !fuzz_E) {
        printf("Usage: fpng_b200_test [-s] [-u] [-a] [-c] [-f] [-e] [-E] [-n trials] [-o out.png] filename.png\n");
        return EXIT_FAILURE;
    }
    fpng::fpng_init();
    if (!fpngb_is_initialized()) { fprintf(stderr, "no CUDA device: this implementation has no CPU path\n"); return EXIT_FAILURE; }
    fpngb_set_fallback_decoder(fallback_cb, nullptr);
    const uint32_t flags = (slower ? fpng::FPNG_ENCODE_SLOWER : 0) | (uncompressed ? fpng::FPNG_FORCE_UNCOMPRESSED : 0);

    if (fuzz_E) {                                                            // src/fpng_test.cpp:617-682
        void* s = fuzzgen_dims_open();
        for (uint32_t t = 0; t < trials; t++) {
            uint32_t w, h, c; fuzzgen_dims_next(s, &w, &h, &c);
            bytes img((size_t)w * h * c);
            fuzzgen_dims_fill(s, img.data(), (uint64_t)w * h, c);
            printf("Testing %ux%u %u\n", w, h, c);
            size_t sz = 0;
            bytes png;
            if (!fpng::fpng_encode_image_to_memory(img.data(), w, h, c, png, flags)) { fprintf(stderr, "fpng_encode_image_to_memory() failed!\n"); return EXIT_FAILURE; }
            sz = png.size();
            printf("fpng size: %u\n", (uint32_t)sz);
            bytes dec; uint32_t dw, dh, dc;
            if (fpng::fpng_decode_memory(png.data(), (uint32_t)png.size(), dec, dw, dh, dc, c) != fpng::FPNG_DECODE_SUCCESS || dw != w || dh != h || dc != c || dec != img) {
                fprintf(stderr, "Decoded image failed verification\n"); return EXIT_FAILURE; }
        }
        fuzzgen_dims_close(s);
        return EXIT_SUCCESS;
    }

    bytes file;
    if (!read_file(filename, file)) { fprintf(stderr, "Failed reading %s\n", filename); return EXIT_FAILURE; }
    if (decode_only) {                                                       // src/fpng_test.cpp:1092-1114 (with the fallback hook off: the raw fpng status)
        fpngb_set_fallback_decoder(nullptr, nullptr);
        bytes px; uint32_t w, h, c;
        const int st = fpng::fpng_decode_memory(file.data(), (uint32_t)file.size(), px, w, h, c, 4);
        if (st != fpng::FPNG_DECODE_SUCCESS) { fprintf(stderr, "fpng_decode_memory() failed with error %i\n", st); return EXIT_FAILURE; }
        printf("fpng_decode_memory() succeeded: %ux%u, %u channels\n", w, h, c);
        return EXIT_SUCCESS;
    }

    // source image: any 8-bit PNG, through fpng::fpng_decode_memory -- fpng files on the GPU, others through the fallback hook
    bytes rgba; uint32_t w = 0, h = 0, file_chans = 0;
    const int lst = fpng::fpng_decode_memory(file.data(), (uint32_t)file.size(), rgba, w, h, file_chans, 4);
    if (lst != fpng::FPNG_DECODE_SUCCESS) { fprintf(stderr, "Failed loading %s (status %d)\n", filename, lst); return EXIT_FAILURE; }
    if (swizzle) for (size_t p = 0; p < (size_t)w * h; p++) rgba[p * 4 + 3] = rgba[p * 4 + 1];
    bool has_alpha = false;                                                  // src/fpng_test.cpp:1156-1166
    for (size_t p = 0; p < (size_t)w * h && !has_alpha; p++) has_alpha = rgba[p * 4 + 3] < 255;
    const uint32_t chans = has_alpha ? 4 : 3;
    bytes img((size_t)w * h * chans);
    for (size_t p = 0; p < (size_t)w * h; p++) for (uint32_t j = 0; j < chans; j++) img[p * chans + j] = rgba[p * 4 + j];
    const double total_px = (double)w * h;
    if (!csv) printf("Filename: %s\nDimensions: %ux%u, Has Alpha: %u, Total Pixels: %u, bytes: %u (%f MB)\n", filename, w, h, has_alpha, (uint32_t)total_px,
                     (uint32_t)img.size(), img.size() / (1024.0 * 1024.0));

    if (fuzz_e) {                                                            // src/fpng_test.cpp:381-615
        for (uint32_t t = 0; t < trials; t++) {
            bytes buf = img;
            const int fam = fuzzgen_mutate(t, buf.data(), (uint32_t)buf.size(), chans);
            size_t sz = 0;
            if (!roundtrip_ok(buf, w, h, chans, flags, &sz)) return EXIT_FAILURE;
            static const char* names[6] = { "color fill runs", "color corrupt runs", "fill runs", "corrupt runs", "full random", "bits flipped" };
            printf("%u, %s\nfpng size: %u\n", t, names[fam], (uint32_t)sz);
        }
        return EXIT_SUCCESS;
    }

    // encode: best of 3 (src/fpng_test.cpp:1181-1209)
    bytes png; double enc_best = 1e9;
    for (int i = 0; i < 3; i++) {
        const double t0 = now_s();
        if (!fpng::fpng_encode_image_to_memory(img.data(), w, h, chans, png, flags)) { fprintf(stderr, "fpng_encode_image_to_memory() failed!\n"); return EXIT_FAILURE; }
        enc_best = std::min(enc_best, now_s() - t0);
    }
    { FILE* f = fopen(out_name, "wb"); if (!f || fwrite(png.data(), 1, png.size(), f) != png.size()) { fprintf(stderr, "Failed writing %s\n", out_name); return EXIT_FAILURE; } fclose(f); }
    // decode: best of 5, verify (src/fpng_test.cpp:1237-1273)
    double dec_best = 1e9; bytes dec; uint32_t dw, dh, dc;
    for (int i = 0; i < 5; i++) {
        const double t0 = now_s();
        const int st = fpng::fpng_decode_memory(png.data(), (uint32_t)png.size(), dec, dw, dh, dc, chans);
        dec_best = std::min(dec_best, now_s() - t0);
        if (st != fpng::FPNG_DECODE_SUCCESS || dw != w || dh != h || dc != chans || dec != img) { fprintf(stderr, "fpng decode verification failed (status %d)\n", st); return EXIT_FAILURE; }
    }
    // channel conversions (src/fpng_test.cpp:1276-1327)
    {
        const uint32_t other = chans == 3 ? 4 : 3;
        if (fpng::fpng_decode_memory(png.data(), (uint32_t)png.size(), dec, dw, dh, dc, other) != fpng::FPNG_DECODE_SUCCESS) { fprintf(stderr, "conversion decode failed\n"); return EXIT_FAILURE; }
        for (size_t p = 0; p < (size_t)w * h; p++) {
            for (uint32_t j = 0; j < 3; j++) if (dec[p * other + j] != img[p * chans + j]) { fprintf(stderr, "conversion verification failed\n"); return EXIT_FAILURE; }
            if (other == 4 && dec[p * 4 + 3] != 0xFF) { fprintf(stderr, "conversion alpha verification failed\n"); return EXIT_FAILURE; }
        }
    }
    // independent decoder (checks IDAT CRC-32 and Adler-32 like lodepng does, src/fpng_test.cpp:1330-1363)
    { bytes gen; uint32_t gw, gh, gc;
      if (!general_png_decode(png.data(), png.size(), chans, gen, gw, gh, gc) || gw != w || gh != h || gen != img) { fprintf(stderr, "independent PNG decoder verification failed\n"); return EXIT_FAILURE; } }
    // the reference's own decoder accepts the file iff fpng_get_info says so
    { uint32_t iw, ih, ic; if (fpng::fpng_get_info(png.data(), (uint32_t)png.size(), iw, ih, ic) != fpng::FPNG_DECODE_SUCCESS || iw != w || ih != h || ic != chans) { fprintf(stderr, "fpng_get_info failed\n"); return EXIT_FAILURE; } }

    const double mib = 1024.0 * 1024.0, smp = total_px / mib;               // the reference reports mebi-pixels (src/fpng_test.cpp:1212)
    if (!csv) {
        printf("** Encoding:\nFPNG:    %3.6f secs, %u bytes, %4.3f MB, %4.3f MP/sec\n", enc_best, (uint32_t)png.size(), png.size() / mib, smp / enc_best);
        printf("** Decoding:\nFPNG:    %3.6f secs, %4.3f MP/sec\n", dec_best, smp / dec_best);
        printf("Verified: fpng decode (%u ch), %u<->%u channel conversion, independent zlib-based PNG decoder (CRC-32 + Adler-32 checked)\n", chans, chans, chans == 3 ? 4 : 3);
    } else {
        printf("%s, %u, %u, %u,    %f, %f, %f, %4.3f, %4.3f\n", filename, w, h, chans, enc_best, png.size() / mib, dec_best, smp / enc_best, smp / dec_best);
    }
    return EXIT_SUCCESS;
}
