// fpng_b200/csrc/host_api.cu -- host runtime + C ABI (include/fpng_b200.h) of the B200 fpng hot path.
//
// The host side owns: the CUDA context/stream, the static (1-pass) code books derived from the format's
// pre-serialised block headers, a grow-only device workspace, the 58-byte container header template, and the
// H2D/D2H plumbing for host-buffer callers.  All pixel/bit work runs in the kernels of encode_kernels.cu,
// checksum_kernels.cu, huffman_kernels.cu and decode_kernels.cu.  There is no CPU fallback: without a CUDA
// device every entry point fails with FPNGB_ERR_NO_DEVICE.
#include "../../include/fpng_b200.h"
#include "kernels.cuh"
#include "row_walk.cuh"
#include "static_tables.h"

#include <atomic>
#include <mutex>
#include <new>
#include <string.h>
#include <vector>

namespace fpngb {

static std::atomic<uint64_t> g_launches{0};
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------------------
// Static code books: parse "HLIT HDIST HCLEN, code-length code, RLE'd code sizes" exactly as an inflater does
// (RFC 1951 3.2.7; the reference's own reader is fpng.cpp:1954-2076) and assign canonical codes (fpng.cpp:701-708).
// ---------------------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p; uint32_t nbits; uint32_t pos = 0; bool bad = false;
    uint32_t get(int n)
    {
        uint32_t v = 0;
        for (int i = 0; i < n; i++) {
            if (pos >= nbits) { bad = true; return 0; }
            v |= (uint32_t)((p[pos >> 3] >> (pos & 7)) & 1) << i;
            pos++;
        }
        return v;
    }
};

static uint32_t reverse_bits(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) { r = (r << 1) | ((v >> i) & 1); } return r; }

static void canonical_codes(const uint8_t* sizes, int n, uint16_t* codes)
{
    uint32_t count[16] = {0}, next[16] = {0};
    for (int i = 0; i < n; i++) count[sizes[i]]++;
    count[0] = 0;
    for (int l = 1; l < 16; l++) next[l] = (next[l - 1] + count[l - 1]) << 1;
    for (int i = 0; i < n; i++) codes[i] = sizes[i] ? (uint16_t)reverse_bits(next[sizes[i]]++, sizes[i]) : 0;
}

static bool read_code_sizes(BitReader& br, uint8_t* lit_sizes /*288*/, uint8_t* dist_sizes /*32*/)
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const uint32_t nlit = br.get(5) + 257, ndist = br.get(5) + 1, nclen = br.get(4) + 4;
    uint8_t cl[19] = {0};
    for (uint32_t i = 0; i < nclen; i++) cl[order[i]] = (uint8_t)br.get(3);
    uint16_t clcodes[19];
    canonical_codes(cl, 19, clcodes);
    uint8_t all[288 + 32] = {0};
    uint32_t cur = 0;
    while (cur < nlit + ndist && !br.bad) {
        // decode one code-length symbol bit by bit (tiny table: linear search is fine at init time)
        uint32_t code = 0; int len = 0, sym = -1;
        while (sym < 0 && len < 8) {
            code |= br.get(1) << len; len++;
            for (int s = 0; s < 19; s++) if (cl[s] == len && clcodes[s] == code) { sym = s; break; }
        }
        if (sym < 0) return false;
        if (sym < 16) { all[cur++] = (uint8_t)sym; continue; }
        uint32_t rep, val = 0;
        if (sym == 16) { rep = br.get(2) + 3; if (!cur) return false; val = all[cur - 1]; }
        else if (sym == 17) rep = br.get(3) + 3;
        else rep = br.get(7) + 11;
        if (cur + rep > nlit + ndist) return false;
        while (rep--) all[cur++] = (uint8_t)val;
    }
    if (br.bad) return false;
    memset(lit_sizes, 0, 288); memset(dist_sizes, 0, 32);
    memcpy(lit_sizes, all, nlit); memcpy(dist_sizes, all + nlit, ndist);
    return true;
}

// Fill the kernel-facing tables of a code book from lit/len code sizes + codes.
void finish_codebook(CodeBook& cb, const uint8_t* sizes, const uint16_t* codes, uint32_t chans)
{
    for (int v = 0; v < 256; v++) { cb.lit[v] = codes[v] | ((uint32_t)sizes[v] << 16); cb.lit_size[v] = sizes[v]; }
    memset(cb.match, 0, sizeof cb.match); memset(cb.match_bits, 0, sizeof cb.match_bits);
    const uint32_t M = max_match_pixels(chans);
    for (uint32_t n = 1; n <= M; n++) {
        uint32_t sym, xb, xv;
        deflate_len_code(n * chans, sym, xb, xv);
        const uint32_t s = sizes[sym], total = s + xb + 1;          // +1: the 1-bit distance code, always 0 (fpng.cpp:1135)
        cb.match[n] = (codes[sym] | (xv << s)) | (total << 24);
        cb.match_bits[n] = (uint8_t)total;
    }
    cb.eob = codes[256] | ((uint32_t)sizes[256] << 16);
    memcpy(cb.sym_size, sizes, 288);
}

static bool build_static_book(CodeBook& cb, const uint8_t* hdr, uint32_t nbytes, uint32_t tail, uint32_t tail_bits, uint32_t chans)
{
    memset(&cb, 0, sizeof cb);
    memcpy(cb.hdr, hdr, nbytes);
    cb.hdr[nbytes] = (uint8_t)tail;
    cb.hdr_bits = nbytes * 8 + tail_bits;
    BitReader br{cb.hdr, cb.hdr_bits};
    br.pos = 16;
    if (br.get(1) != 1 || br.get(2) != 2) return false;
    uint8_t lit_sizes[288], dist_sizes[32];
    if (!read_code_sizes(br, lit_sizes, dist_sizes) || br.pos != cb.hdr_bits) return false;
    if (dist_sizes[chans - 1] != 1) return false;
    uint16_t codes[288];
    canonical_codes(lit_sizes, 288, codes);
    const uint32_t M = max_match_pixels(chans);
    for (uint32_t n = 1; n <= M; n++) { uint32_t s, xb, xv; deflate_len_code(n * chans, s, xb, xv); if (!lit_sizes[s]) return false; }
    for (int v = 0; v <= 256; v++) if (!lit_sizes[v]) return false;
    finish_codebook(cb, lit_sizes, codes, chans);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------
struct Buffer {
    void* p = nullptr; size_t cap = 0; bool pinned = false;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
        n = (n + (1u << 20)) & ~((size_t)(1u << 20) - 1);
        FPNGB_CUDA_OK(pinned ? cudaMallocHost(&p, n) : cudaMalloc(&p, n));
        cap = n;
        return 0;
    }
};

struct Context {
    int device = -1;
    cudaStream_t stream = nullptr;
    CodeBook* d_static_books = nullptr;          // [0] RGB, [1] RGBA
    CodeBook h_static_books[2];
    Buffer ws;                                    // kernel workspace (row tables, image state, histograms, books)
    Buffer dev_in, dev_out;                       // staging for the *_host entry points
    Buffer pin_small;                             // pinned scratch for sizes / status words
    std::mutex mu;                                // the *_host entry points and the workspace are serialised
    bool ready = false;
};

static Context g_ctx;
static std::mutex g_init_mu;

struct Workspace {
    uint32_t* row_bits; uint2* row_adler; unsigned long long* row_ofs; ImageState* st; uint32_t* hist; CodeBook* books;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int carve_workspace(Context& c, uint32_t n, uint32_t h, bool two_pass, Workspace& w)
{
    const size_t rows = (size_t)n * h;
    size_t o = 0;
    const size_t o_bits = o; o = align_up(o + rows * 4, 256);
    const size_t o_adl = o; o = align_up(o + rows * 8, 256);
    const size_t o_ofs = o; o = align_up(o + rows * 8, 256);
    const size_t o_st = o; o = align_up(o + (size_t)n * sizeof(ImageState), 256);
    const size_t o_hist = o; o = align_up(o + (two_pass ? (size_t)n * 288 * 4 : 0), 256);
    const size_t o_books = o; o = align_up(o + (two_pass ? (size_t)n * sizeof(CodeBook) : 0), 256);
    int rc = c.ws.reserve(o);
    if (rc) return rc;
    uint8_t* b = (uint8_t*)c.ws.p;
    w.row_bits = (uint32_t*)(b + o_bits); w.row_adler = (uint2*)(b + o_adl); w.row_ofs = (unsigned long long*)(b + o_ofs);
    w.st = (ImageState*)(b + o_st); w.hist = (uint32_t*)(b + o_hist); w.books = (CodeBook*)(b + o_books);
    return 0;
}

static int pick_load_mode(const void* base, size_t image_stride, uint32_t w, uint32_t chans)
{
    const uintptr_t a = (uintptr_t)base;
    if (chans == 4) {
        if (a % 16 == 0 && image_stride % 16 == 0 && w % 4 == 0) return kLoadVec16;
        if (a % 4 == 0 && image_stride % 4 == 0) return kLoadWords;
        return kLoadBytes;
    }
    if (a % 4 == 0 && image_stride % 4 == 0 && w % 4 == 0) return kLoadWords;
    return kLoadBytes;
}

static bool valid_dims(uint32_t w, uint32_t h, uint32_t chans)
{
    // fpng.cpp:1670-1680; additionally the filtered stream must be addressable with 32 bits like the reference's
    // own uint32 offsets.
    if (w < 1 || h < 1 || (uint64_t)w * h > 0xFFFFFFFFull || w > (1u << 24) || h > (1u << 24)) return false;
    if (chans != 3 && chans != 4) return false;
    if (((uint64_t)w * chans + 1) * h + 1024 > 0xFFFFFFFFull) return false;
    return true;
}

static void make_png_header(uint8_t* hdr, uint32_t w, uint32_t h, uint32_t chans)
{
    // fpng.cpp:1770-1783 (Appendix A of SURVEY.md).  Width/height are written as full big-endian 32-bit values
    // (the reference only writes the low 16 bits, Q1: identical for every dimension < 65536).
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    static const uint8_t fdec[17] = {0, 0, 0, 5, 'f', 'd', 'E', 'C', 82, 36, 147, 227, 0, 0xE5, 0xAB, 0x62, 0x99};
    memset(hdr, 0, kPngHeaderSize);
    memcpy(hdr, sig, 8);
    hdr[11] = 13; memcpy(hdr + 12, "IHDR", 4);
    for (int i = 0; i < 4; i++) { hdr[16 + i] = (uint8_t)(w >> (24 - 8 * i)); hdr[20 + i] = (uint8_t)(h >> (24 - 8 * i)); }
    hdr[24] = 8; hdr[25] = chans == 3 ? 2 : 6;
    const uint32_t c = host_crc32(hdr + 12, 17, 0);
    for (int i = 0; i < 4; i++) hdr[29 + i] = (uint8_t)(c >> (24 - 8 * i));
    memcpy(hdr + 33, fdec, 17);
    memcpy(hdr + 54, "IDAT", 4);            // bytes 50..53 (IDAT length) are patched per image on the device
}

// Enqueue the whole encode pipeline for a device-resident batch.  Caller holds ctx.mu.
static int encode_batch_locked(Context& c, const uint8_t* d_pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h,
                               uint32_t chans, uint32_t flags, uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, cudaStream_t s)
{
    const bool two_pass = (flags & FPNGB_ENCODE_SLOWER) && !(flags & FPNGB_FORCE_UNCOMPRESSED);
    Workspace ws;
    int rc = carve_workspace(c, n, h, two_pass, ws);
    if (rc) return rc;
    const int mode = pick_load_mode(d_pixels, image_stride, w, chans);
    const CodeBook* books = two_pass ? ws.books : c.d_static_books + (chans == 4 ? 1 : 0);
    const uint32_t book_stride = two_pass ? 1u : 0u;

    ScanParams sp{};
    sp.pixels = d_pixels; sp.image_stride = image_stride; sp.w = w; sp.h = h;
    sp.books = books; sp.book_stride = book_stride;
    sp.row_bits = ws.row_bits; sp.row_adler = ws.row_adler; sp.st = ws.st; sp.hist = ws.hist;
    sp.merge_first_unit = (!two_pass && chans == 3) ? 1u : 0u;

    if (two_pass) {
        FPNGB_CUDA_OK(cudaMemsetAsync(ws.hist, 0, (size_t)n * 288 * 4, s));
        launch_scan(sp, n, chans, mode, true, s);
        HuffParams hp{ws.hist, ws.books, chans};
        launch_huffman_build(hp, n, s);
        count_launch(3);
    }
    launch_scan(sp, n, chans, mode, false, s);

    OffsetsParams op{};
    op.row_bits = ws.row_bits; op.row_ofs = ws.row_ofs; op.books = books; op.book_stride = book_stride; op.st = ws.st;
    op.out = d_out; op.out_stride = out_stride; op.sizes = d_sizes; op.w = w; op.h = h; op.chans = chans; op.flags = flags;
    make_png_header(op.png_header, w, h, chans);
    launch_offsets(op, n, s);

    PackParams pp{};
    pp.pixels = d_pixels; pp.image_stride = image_stride; pp.w = w; pp.h = h; pp.books = books; pp.book_stride = book_stride;
    pp.row_ofs = ws.row_ofs; pp.row_adler = ws.row_adler; pp.st = ws.st; pp.out = d_out; pp.out_stride = out_stride;
    launch_pack(pp, n, chans, mode, s);

    AdlerParams ap{ws.row_adler, ws.st, d_out, out_stride, w, h, chans};
    launch_adler_finalize(ap, n, s);

    CrcParams cp{};
    cp.out = d_out; cp.out_stride = out_stride; cp.st = ws.st;
    cp.max_tiles = crc_ctas_for(max_encoded_size(w, h, chans)); cp.msg_start = kPngHeaderSize - 4; cp.init_xor = 0xFFFFFFFFu;
    launch_crc(cp, n, s);
    count_launch(5);
    FPNGB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fpngb

using namespace fpngb;

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int fpngb_init(int device)
{
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_ctx.ready) return FPNGB_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FPNGB_ERR_NO_DEVICE;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return FPNGB_ERR_NO_DEVICE; }
    if (device >= count) return FPNGB_ERR_INVALID_ARG;
    FPNGB_CUDA_OK(cudaSetDevice(device));
    Context& c = g_ctx;
    c.device = device;
    FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    if (!build_static_book(c.h_static_books[0], kStaticHdrRGB, sizeof kStaticHdrRGB, kStaticHdrRGBTail, kStaticHdrRGBTailBits, 3)) return FPNGB_ERR_INTERNAL;
    if (!build_static_book(c.h_static_books[1], kStaticHdrRGBA, sizeof kStaticHdrRGBA, kStaticHdrRGBATail, kStaticHdrRGBATailBits, 4)) return FPNGB_ERR_INTERNAL;
    FPNGB_CUDA_OK(cudaMalloc(&c.d_static_books, 2 * sizeof(CodeBook)));
    FPNGB_CUDA_OK(cudaMemcpy(c.d_static_books, c.h_static_books, 2 * sizeof(CodeBook), cudaMemcpyHostToDevice));
    int rc = checksum_tables_init();
    if (rc) return rc;
    c.pin_small.pinned = true;
    rc = c.pin_small.reserve(1 << 16);
    if (rc) return rc;
    c.ready = true;
    return FPNGB_OK;
}

int fpngb_is_initialized(void) { return g_ctx.ready ? 1 : 0; }
const char* fpngb_version(void) { return "fpng_b200 0.1 (sm_100a; format-compatible with fpng 1.0.6)"; }
uint64_t fpngb_launch_count(void) { return g_launches.load(); }

size_t fpngb_max_encoded_size(uint32_t w, uint32_t h, uint32_t chans) { return max_encoded_size(w, h, chans); }

// exposes the static code books to the tests (sizes[288], codes[288]); not part of the reference surface
FPNGB_API int fpngb_debug_static_table(uint32_t chans, uint8_t* sizes, uint16_t* codes, uint32_t* hdr_bits)
{
    CodeBook cb;
    const bool ok = chans == 3 ? build_static_book(cb, kStaticHdrRGB, sizeof kStaticHdrRGB, kStaticHdrRGBTail, kStaticHdrRGBTailBits, 3)
                               : build_static_book(cb, kStaticHdrRGBA, sizeof kStaticHdrRGBA, kStaticHdrRGBATail, kStaticHdrRGBATailBits, 4);
    if (!ok) return FPNGB_ERR_INTERNAL;
    memcpy(sizes, cb.sym_size, 288);
    for (int i = 0; i < 256; i++) codes[i] = (uint16_t)(cb.lit[i] & 0xFFFF);
    for (int i = 256; i < 288; i++) codes[i] = 0;
    codes[256] = (uint16_t)(cb.eob & 0xFFFF);
    *hdr_bits = cb.hdr_bits;
    return FPNGB_OK;
}

int fpngb_encode_batch_device(const void* d_pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h, uint32_t chans,
                              uint32_t flags, void* d_out, size_t out_stride, uint32_t* d_sizes, void* stream)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!d_pixels || !d_out || !d_sizes || n == 0 || !valid_dims(w, h, chans)) return FPNGB_ERR_INVALID_ARG;
    if (image_stride < (size_t)w * h * chans) return FPNGB_ERR_INVALID_ARG;
    if (out_stride < max_encoded_size(w, h, chans)) return FPNGB_ERR_BUFFER_TOO_SMALL;
    if (out_stride % 16 || (uintptr_t)d_out % 16) return FPNGB_ERR_ALIGNMENT;
    if (n > 65535) return FPNGB_ERR_INVALID_ARG;
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
    return encode_batch_locked(c, (const uint8_t*)d_pixels, image_stride, n, w, h, chans, flags, (uint8_t*)d_out, out_stride, d_sizes, s);
}

int fpngb_encode_host(const void* pixels, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags, void* out, size_t out_cap, size_t* out_size)
{
    if (out_size) *out_size = 0;
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!pixels || !out || !out_size || !valid_dims(w, h, chans)) return FPNGB_ERR_INVALID_ARG;
    const size_t cap = max_encoded_size(w, h, chans);
    if (out_cap < cap) return FPNGB_ERR_BUFFER_TOO_SMALL;
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    const size_t in_bytes = (size_t)w * h * chans, stride = align_up(cap, 16);
    int rc = c.dev_in.reserve(in_bytes + 64); if (rc) return rc;
    rc = c.dev_out.reserve(stride + 64); if (rc) return rc;
    uint32_t* d_size = (uint32_t*)((uint8_t*)c.dev_out.p + stride);
    cudaStream_t s = c.stream;
    FPNGB_CUDA_OK(cudaMemcpyAsync(c.dev_in.p, pixels, in_bytes, cudaMemcpyHostToDevice, s));
    rc = encode_batch_locked(c, (const uint8_t*)c.dev_in.p, in_bytes, 1, w, h, chans, flags, (uint8_t*)c.dev_out.p, stride, d_size, s);
    if (rc) return rc;
    uint32_t* h_size = (uint32_t*)c.pin_small.p;
    FPNGB_CUDA_OK(cudaMemcpyAsync(h_size, d_size, 4, cudaMemcpyDeviceToHost, s));
    FPNGB_CUDA_OK(cudaStreamSynchronize(s));
    const size_t fsize = *h_size;
    if (fsize < kPngHeaderSize + kPngTrailerSize || fsize > cap) return FPNGB_ERR_INTERNAL;
    FPNGB_CUDA_OK(cudaMemcpyAsync(out, c.dev_out.p, fsize, cudaMemcpyDeviceToHost, s));
    FPNGB_CUDA_OK(cudaStreamSynchronize(s));
    *out_size = fsize;
    return FPNGB_OK;
}

void* fpngb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void fpngb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
