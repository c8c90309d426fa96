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
#include "runtime.h"

#include <atomic>
#include <mutex>
#include <new>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace fpngb {

static std::atomic<uint64_t> g_launches{0};

static bool g_pack_crc = true;                  // fpngb_debug_pack_crc(): the 16-pixel pack kernel computes scanline CRCs (no CRC pass over the file)
static bool g_inline_crc = true;                // fpngb_debug_inline_crc(): the single-pass encoder computes the IDAT CRC from in-kernel partials
static bool g_crc_stream = true;                // fpngb_debug_crc_stream(0): first-generation (tile-staging) IDAT CRC kernel
static int g_fused_mode = -1;                   // fpngb_debug_use_fused(): 1 single-pass encoder, 0 two-kernel encoder, -1 environment (FPNGB_FUSED)
static bool g_crc_overlap = false;              // fpngb_debug_crc_overlap(1): chunked batches with the CRC kernel on a side stream (measured slower: profiles/README.md)
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
    // fpng decoders (fpng.cpp:2007, 2058-2074) accept literal/length code sizes up to 12 only, and the pack kernels
    // hard-code the distance code as the single bit 0: exactly the table the trainer emits (fpng.cpp:1096-1098) --
    // 1-bit codes at distance symbols chans-1 (code 0) and optionally chans (code 1), nothing else.
    for (int v = 0; v < 288; v++) if (lit_sizes[v] > 12) return false;
    for (uint32_t d = 0; d < 32; d++) {
        const bool allowed = d == chans - 1 || d == chans;
        if (dist_sizes[d] != 0 && !(allowed && dist_sizes[d] == 1)) return false;
    }
    if (dist_sizes[chans - 1] != 1) return false;
    uint16_t codes[288];
    canonical_codes(lit_sizes, 288, codes);
    const uint32_t M = max_match_pixels(chans);
    for (uint32_t n = 1; n <= M; n++) { uint32_t s, xb, xv; deflate_len_code(n * chans, s, xb, xv); if (!lit_sizes[s]) return false; }
    for (int v = 0; v <= 256; v++) if (!lit_sizes[v]) return false;
    finish_codebook(cb, lit_sizes, codes, chans);
    // can "size(len sym 258) + 1 > sum of four literal sizes" (fpng.cpp:1520-1528) ever hold under this table?
    uint32_t min_lit = 255;
    for (int v = 0; v < 256; v++) min_lit = lit_sizes[v] < min_lit ? lit_sizes[v] : min_lit;
    cb.lit1_rule = (chans == 4 && cb.match_bits[1] > 4u * min_lit) ? 1 : 0;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------
static Context g_ctx;
static std::mutex g_init_mu;
Context& context() { return g_ctx; }

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
enum ProfSlot { kProfHist = 0, kProfHuff, kProfScan, kProfOffsets, kProfFused, kProfFinish, kProfPack, kProfAdler, kProfCrc, kProfSlots };
struct ProfSet { cudaEvent_t ev[kProfSlots + 1]; cudaEvent_t crc_start; bool used[kProfSlots]; bool crc_side; uint64_t call; };
static uint64_t g_prof_call = 0;                // encode calls seen while profiling (a chunked call owns several ProfSets)
static bool g_profile = false;
static std::vector<ProfSet*> g_prof_sets;      // one per encode call since the last read
static std::vector<ProfSet*> g_prof_free;

static ProfSet* prof_begin(cudaStream_t s)
{
    if (!g_profile) return nullptr;
    ProfSet* ps;
    if (!g_prof_free.empty()) { ps = g_prof_free.back(); g_prof_free.pop_back(); }
    else {
        ps = new ProfSet();
        for (int i = 0; i <= kProfSlots; i++) if (cudaEventCreate(&ps->ev[i]) != cudaSuccess) { delete ps; return nullptr; }
        if (cudaEventCreate(&ps->crc_start) != cudaSuccess) { delete ps; return nullptr; }
    }
    ps->crc_side = false; ps->call = g_prof_call;
    for (int i = 0; i < kProfSlots; i++) ps->used[i] = false;
    cudaEventRecord(ps->ev[0], s);
    g_prof_sets.push_back(ps);
    return ps;
}
// records the end of slot `slot`; slots must be closed in increasing order, skipped slots stay unused
static void prof_mark(ProfSet* ps, int slot, cudaStream_t s)
{
    if (!ps) return;
    cudaEventRecord(ps->ev[slot + 1], s);
    ps->used[slot] = true;
}

struct Workspace {
    uint32_t* row_bits; uint2* row_adler; unsigned long long* row_ofs; ImageState* st; uint32_t* hist; CodeBook* books;
    uint2* lane_ofs; uint32_t lane_ofs_pitch;
    void* fused_desc;
    uint32_t* row_crc;
};


static int carve_workspace(Context& c, uint32_t n, uint32_t h, uint32_t width, bool two_pass, Workspace& w, bool fused = false)
{
    const size_t rows = (size_t)n * h;
    const uint32_t lane_pitch = ((width + 511u) / 512u) * 32u;       // one uint2 per 16-pixel group, whole warp steps
    size_t o = 0;
    const size_t o_bits = o; o = align_up(o + rows * 4, 256);
    const size_t o_adl = o; o = align_up(o + rows * 8, 256);
    const size_t o_ofs = o; o = align_up(o + rows * 8, 256);
    const size_t o_st = o; o = align_up(o + (size_t)n * sizeof(ImageState), 256);
    const size_t o_hist = o; o = align_up(o + (two_pass ? (size_t)n * 288 * 4 : 0), 256);
    const size_t o_books = o; o = align_up(o + (two_pass ? (size_t)n * sizeof(CodeBook) : 0), 256);
    const size_t o_lane = o; o = align_up(o + (fused ? 0 : rows * lane_pitch * 8), 256);
    const size_t o_desc = o; o = align_up(o + (fused ? fused_desc_bytes(n, width, h) : 0), 256);
    const size_t o_rcrc = o; o = align_up(o + (fused ? 0 : rows * 4), 256);
    int rc = c.ws.reserve(o);
    if (rc) return rc;
    uint8_t* b = (uint8_t*)c.ws.p;
    w.row_crc = (uint32_t*)(b + o_rcrc);
    w.row_bits = (uint32_t*)(b + o_bits); w.row_adler = (uint2*)(b + o_adl); w.row_ofs = (unsigned long long*)(b + o_ofs);
    w.st = (ImageState*)(b + o_st); w.hist = (uint32_t*)(b + o_hist); w.books = (CodeBook*)(b + o_books);
    w.lane_ofs = (uint2*)(b + o_lane); w.lane_ofs_pitch = lane_pitch;
    w.fused_desc = b + o_desc;
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
//
// Two encoders share this entry:
//  * the two-kernel encoder (scan -> offsets -> pack -> adler -> crc), the default.  Optionally (fpngb_debug_crc_overlap) large
//    batches are cut into chunks and the CRC kernel of chunk k runs on a side stream underneath the scan and pack kernels of
//    chunk k + 1 -- measured SLOWER on B200 (C2: 2.79 vs 2.60 ms: the kernels slow each other down more than the overlap
//    gains), so it is off by default;
//  * the single-pass encoder (encode_fused.cu), selected with FPNGB_FUSED=1 / fpngb_debug_use_fused(1): reads the pixels once
//    and writes the file once, byte-identical output, measured slower on B200 (profiles/README.md) because the path is bound by
//    instruction issue, not by HBM.
static int encode_batch_locked(Context& c, const uint8_t* d_pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h,
                               uint32_t chans, uint32_t flags, uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, cudaStream_t s)
{
    const bool two_pass = (flags & FPNGB_ENCODE_SLOWER) && !(flags & FPNGB_FORCE_UNCOMPRESSED);
    // second-generation kernels (16 pixels per lane, coalesced 128-bit loads) whenever every scanline is 16-byte aligned;
    // FPNGB_FORCE_GENERIC=1 keeps the generic kernels (tests compare both)
    static const bool force_generic = getenv("FPNGB_FORCE_GENERIC") && atoi(getenv("FPNGB_FORCE_GENERIC")) != 0;
    static const bool env_fused = getenv("FPNGB_FUSED") && atoi(getenv("FPNGB_FUSED")) != 0;
    // RGBA 1-pass under a table where a one-pixel match can lose against four literals: only the generic kernels
    // implement the reference's check (fpng.cpp:1520-1528); it cannot fire with the shipped table
    const uint32_t lit1_rule = (!two_pass && chans == 4 && c.h_static_books[1].lit1_rule) ? 1u : 0u;
    // the 16-pixel kernels take any scanline alignment (staged-tile loader for 16-byte aligned scanlines, direct realigning
    // loads otherwise); the generic 4-pixel kernels remain for the trained-table rule and for A/B tests
    const bool v2 = !force_generic && !lit1_rule;
    const bool want_fused = g_fused_mode == 1 || (g_fused_mode < 0 && env_fused);
    const bool fused = want_fused && !force_generic && !lit1_rule && !(flags & FPNGB_FORCE_UNCOMPRESSED) &&
                       fused_eligible(d_pixels, image_stride, w, h, chans, n);
    Workspace ws;
    int rc = carve_workspace(c, n, h, w, two_pass, ws, fused);
    if (rc) return rc;
    rc = c.ws_acquire(s);
    if (rc) return rc;
    const int mode = pick_load_mode(d_pixels, image_stride, w, chans);
    const CodeBook* books0 = two_pass ? ws.books : c.d_static_books + (chans == 4 ? 1 : 0);
    const uint32_t book_stride = two_pass ? 1u : 0u;
    const uint32_t merge_first_unit = (!two_pass && chans == 3) ? 1u : 0u;
    uint8_t png_header[kPngHeaderSize];
    make_png_header(png_header, w, h, chans);
    if (g_profile) g_prof_call++;

    // chunks: the CRC of one chunk overlaps the scan/pack of the next.  Small batches (and the single-pass encoder) run as one chunk.
    const uint32_t nchunks = (!fused && g_crc_overlap && n >= 16) ? 4u : 1u;
    if (nchunks > 1) {
        if (!c.side) FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
        for (int i = 0; i < 5; i++) if (!c.side_ev[i]) FPNGB_CUDA_OK(cudaEventCreateWithFlags(&c.side_ev[i], cudaEventDisableTiming));
    }
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t i0 = (uint32_t)((uint64_t)n * k / nchunks), i1 = (uint32_t)((uint64_t)n * (k + 1) / nchunks), cnt = i1 - i0;
        if (!cnt) continue;
        const uint8_t* px = d_pixels + (size_t)i0 * image_stride;
        uint8_t* out = d_out + (size_t)i0 * out_stride;
        const CodeBook* books = books0 + (size_t)i0 * book_stride;
        const size_t r0 = (size_t)i0 * h;

        ScanParams sp{};
        sp.pixels = px; sp.image_stride = image_stride; sp.w = w; sp.h = h;
        sp.books = books; sp.book_stride = book_stride;
        sp.row_bits = ws.row_bits + r0; sp.row_adler = ws.row_adler + r0; sp.st = ws.st + i0; sp.hist = ws.hist + (size_t)i0 * 288;
        sp.merge_first_unit = merge_first_unit;
        sp.lane_ofs = ws.lane_ofs + r0 * ws.lane_ofs_pitch; sp.lane_ofs_pitch = ws.lane_ofs_pitch;
        sp.lit1_rule = lit1_rule;

        ProfSet* ps = prof_begin(s);
        if (two_pass) {
            FPNGB_CUDA_OK(cudaMemsetAsync(sp.hist, 0, (size_t)cnt * 288 * 4, s));
            if (v2) launch_hist16(sp, cnt, chans, s); else launch_scan(sp, cnt, chans, mode, true, s);
            prof_mark(ps, kProfHist, s);
            HuffParams hp{sp.hist, ws.books + i0, chans, 0u};
            launch_huffman_build(hp, cnt, s);
            prof_mark(ps, kProfHuff, s);
            count_launch(2);
        }
        PackParams pp{};
        pp.pixels = px; pp.image_stride = image_stride; pp.w = w; pp.h = h; pp.books = books; pp.book_stride = book_stride;
        pp.row_ofs = ws.row_ofs + r0; pp.row_bits = sp.row_bits; pp.lane_ofs = sp.lane_ofs; pp.lane_ofs_pitch = ws.lane_ofs_pitch;
        pp.row_adler = sp.row_adler; pp.st = sp.st; pp.out = out; pp.out_stride = out_stride;
        pp.lit1_rule = lit1_rule;
        // the 16-pixel pack kernel computes each scanline's CRC while its code words are staged; the file is not read back
        const bool pack_crc = v2 && !fused && g_pack_crc;
        if (pack_crc) { pp.row_crc = ws.row_crc + r0; pp.crc_f128b = c.crc_f128b; pp.crc_lane_mul = c.crc_lane_mul; }
        if (fused) {
            cudaEvent_t mid = ps ? ps->ev[kProfFused + 1] : nullptr;
            rc = launch_encode_fused(px, image_stride, cnt, w, h, chans, flags, books, book_stride, sp.row_adler, sp.st, ws.fused_desc,
                                     out, out_stride, d_sizes + i0, png_header, merge_first_unit, s, mid, g_inline_crc);
            if (rc) return rc;
            if (ps) ps->used[kProfFused] = true;
            prof_mark(ps, kProfFinish, s);
            pp.stored_only = 1u;
            launch_pack16(pp, cnt, chans, s);                                   // stored-block images only (fpng.cpp:1728-1758)
            prof_mark(ps, kProfPack, s);
            count_launch(3);
        } else {
            if (v2) launch_scan16(sp, cnt, chans, s); else launch_scan(sp, cnt, chans, mode, false, s);
            prof_mark(ps, kProfScan, s);
            OffsetsParams op{};
            op.row_bits = sp.row_bits; op.row_ofs = ws.row_ofs + r0; op.books = books; op.book_stride = book_stride; op.st = sp.st;
            op.out = out; op.out_stride = out_stride; op.sizes = d_sizes + i0; op.w = w; op.h = h; op.chans = chans; op.flags = flags;
            memcpy(op.png_header, png_header, kPngHeaderSize);
            launch_offsets(op, cnt, s);
            prof_mark(ps, kProfOffsets, s);
            if (v2) launch_pack16(pp, cnt, chans, s); else launch_pack(pp, cnt, chans, mode, s);
            prof_mark(ps, kProfPack, s);
            count_launch(3);
        }
        AdlerParams ap{sp.row_adler, sp.st, out, out_stride, w, h, chans};
        launch_adler_finalize(ap, cnt, s);
        prof_mark(ps, kProfAdler, s);

        CrcParams cp{};
        cp.out = out; cp.out_stride = out_stride; cp.st = sp.st;
        cp.max_tiles = crc_ctas_for(max_encoded_size(w, h, chans)); cp.msg_start = kPngHeaderSize - 4; cp.init_xor = 0xFFFFFFFFu;
        cudaStream_t cs = s;
        if (nchunks > 1) {                                                    // this chunk's CRC goes to the side stream
            FPNGB_CUDA_OK(cudaEventRecord(c.side_ev[k], s));
            FPNGB_CUDA_OK(cudaStreamWaitEvent(c.side, c.side_ev[k], 0));
            cs = c.side;
            if (ps) { cudaEventRecord(ps->crc_start, cs); ps->crc_side = true; }
        }
        if (fused && g_inline_crc) {
            launch_fused_crc(ws.fused_desc, cnt, w, h, books, book_stride, sp.st, out, out_stride, cs);   // combine the in-kernel partials
            cp.stored_only = 1u;                                              // the file-reading CRC kernel is only needed for stored-block images
            count_launch(1);
        }
        if (pack_crc) {
            RowCrcParams rp{pp.row_crc, pp.row_ofs, pp.row_bits, books, book_stride, sp.st, out, out_stride, h};
            launch_row_crc_combine(rp, cnt, cs);
            cp.stored_only = 1u;
            count_launch(1);
        }
        if (g_crc_stream) launch_crc_stream(cp, cnt, max_encoded_size(w, h, chans), cs); else launch_crc(cp, cnt, cs);
        prof_mark(ps, kProfCrc, cs);
        count_launch(2);
    }
    if (nchunks > 1) {                                                        // join: the caller's stream continues after the last CRC
        FPNGB_CUDA_OK(cudaEventRecord(c.side_ev[4], c.side));
        FPNGB_CUDA_OK(cudaStreamWaitEvent(s, c.side_ev[4], 0));
    }
    FPNGB_CUDA_OK(cudaGetLastError());
    return c.ws_release(s);
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
    rc = fused_tables_init();
    if (rc) return rc;
    rc = crc_stream_tables_init();
    if (rc) return rc;
    rc = crc_stream_table_ptrs(&c.crc_f128b, &c.crc_lane_mul);
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

// test hook: scanlines per warp of the 16-pixel scan/pack kernels (0 = automatic); not part of the reference surface
FPNGB_API void fpngb_debug_rows_per_warp(uint32_t v) { set_rows_per_warp16(v); }

// test hook: 1 = use the two-kernel (scan + pack) encoder even where the single-pass encoder applies; not part of the reference surface
FPNGB_API void fpngb_debug_disable_fused(int off) { g_fused_mode = off ? 0 : 1; }
FPNGB_API void fpngb_debug_use_fused(int mode) { g_fused_mode = mode; }
FPNGB_API void fpngb_debug_crc_overlap(int on) { g_crc_overlap = on != 0; }
FPNGB_API void fpngb_debug_crc_stream(int on) { g_crc_stream = on != 0; }
FPNGB_API void fpngb_debug_inline_crc(int on) { g_inline_crc = on != 0; }
// test hook: 0 = the two-kernel encoder leaves the IDAT CRC to the file-reading CRC kernel instead of computing it in the pack kernel
FPNGB_API void fpngb_debug_pack_crc(int on) { g_pack_crc = on != 0; }

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
    cudaStream_t s = (cudaStream_t)stream;      // NULL = the CUDA default stream, exactly as the caller passed it
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

// Pipelined host-buffer batch: chunks of the batch flow H2D -> kernels -> (sizes) -> exact-size D2H through three
// slots; chunk k+1 is enqueued before the host waits for chunk k's sizes, so copies overlap the kernels.
int fpngb_encode_batch_host(const void* pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h, uint32_t chans,
                            uint32_t flags, void* out, size_t out_stride, uint32_t* sizes)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!pixels || !out || !sizes || n == 0 || !valid_dims(w, h, chans)) return FPNGB_ERR_INVALID_ARG;
    const size_t in_bytes = (size_t)w * h * chans, cap = max_encoded_size(w, h, chans), dstride = align_up(cap, 16);
    if (image_stride < in_bytes) return FPNGB_ERR_INVALID_ARG;
    if (out_stride < cap) return FPNGB_ERR_BUFFER_TOO_SMALL;
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));

    constexpr int kSlots = 3;
    const size_t target = (size_t)64 << 20;                         // ~64 MiB of pixels per chunk
    uint32_t per_chunk = (uint32_t)(target / in_bytes); if (per_chunk < 1) per_chunk = 1; if (per_chunk > n) per_chunk = n;
    if (per_chunk > 65535u) per_chunk = 65535u;                     // grid.y limit of the batch kernels (tiny images)
    const size_t slot_in = align_up(per_chunk * in_bytes, 256), slot_out = per_chunk * dstride, slot_sz = align_up(per_chunk * 4, 256);
    int rc = c.dev_in.reserve(kSlots * slot_in); if (rc) return rc;
    rc = c.dev_out.reserve(kSlots * (slot_out + slot_sz)); if (rc) return rc;
    rc = c.pin_small.reserve(kSlots * slot_sz); if (rc) return rc;
    if (!c.copy_in) { FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking)); FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking)); }
    EventSet<4 * kSlots> evs;                                       // destroyed on every exit path
    if (!evs.ok) return FPNGB_ERR_INTERNAL;
    cudaEvent_t* ev_in = evs.ev; cudaEvent_t* ev_done = evs.ev + kSlots; cudaEvent_t* ev_sizes = evs.ev + 2 * kSlots; cudaEvent_t* ev_out = evs.ev + 3 * kSlots;
    const uint32_t nchunks = (n + per_chunk - 1) / per_chunk;
    const bool contiguous = image_stride == in_bytes;
    int result = FPNGB_OK;

    auto finish = [&](uint32_t k) -> int {      // exact-size D2H of chunk k's files
        const int sl = k % kSlots;
        const uint32_t first = k * per_chunk, cnt = (first + per_chunk <= n) ? per_chunk : n - first;
        FPNGB_CUDA_OK(cudaEventSynchronize(ev_sizes[sl]));
        const uint32_t* hs = (const uint32_t*)((uint8_t*)c.pin_small.p + sl * slot_sz);
        const uint8_t* dout = (const uint8_t*)c.dev_out.p + sl * (slot_out + slot_sz);
        for (uint32_t i = 0; i < cnt; i++) {
            const uint32_t fs = hs[i];
            if (fs < kPngHeaderSize + kPngTrailerSize || fs > cap) return FPNGB_ERR_INTERNAL;
            sizes[first + i] = fs;
            FPNGB_CUDA_OK(cudaMemcpyAsync((uint8_t*)out + (size_t)(first + i) * out_stride, dout + (size_t)i * dstride, fs, cudaMemcpyDeviceToHost, c.copy_out));
        }
        FPNGB_CUDA_OK(cudaEventRecord(ev_out[sl], c.copy_out));
        return 0;
    };

    for (uint32_t k = 0; k < nchunks && result == FPNGB_OK; k++) {
        const int sl = k % kSlots;
        const uint32_t first = k * per_chunk, cnt = (first + per_chunk <= n) ? per_chunk : n - first;
        if (k >= kSlots) { rc = (int)cudaEventSynchronize(ev_out[sl]); if (rc) { result = 1000 + rc; break; } }   // slot's previous files have left the device
        uint8_t* din = (uint8_t*)c.dev_in.p + sl * slot_in;
        uint8_t* dout = (uint8_t*)c.dev_out.p + sl * (slot_out + slot_sz);
        uint32_t* dsz = (uint32_t*)(dout + slot_out);
        const uint8_t* src = (const uint8_t*)pixels + (size_t)first * image_stride;
        cudaError_t e = contiguous ? cudaMemcpyAsync(din, src, cnt * in_bytes, cudaMemcpyHostToDevice, c.copy_in)
                                   : cudaMemcpy2DAsync(din, in_bytes, src, image_stride, in_bytes, cnt, cudaMemcpyHostToDevice, c.copy_in);
        if (e != cudaSuccess) { result = 1000 + (int)e; break; }
        cudaEventRecord(ev_in[sl], c.copy_in);
        cudaStreamWaitEvent(c.stream, ev_in[sl], 0);
        rc = encode_batch_locked(c, din, in_bytes, cnt, w, h, chans, flags, dout, dstride, dsz, c.stream);
        if (rc) { result = rc; break; }
        cudaEventRecord(ev_done[sl], c.stream);
        cudaStreamWaitEvent(c.copy_out, ev_done[sl], 0);
        cudaMemcpyAsync((uint8_t*)c.pin_small.p + sl * slot_sz, dsz, cnt * 4, cudaMemcpyDeviceToHost, c.copy_out);
        cudaEventRecord(ev_sizes[sl], c.copy_out);
        if (k >= 1) { rc = finish(k - 1); if (rc) { result = rc; break; } }
    }
    if (result == FPNGB_OK) result = finish(nchunks - 1);
    cudaStreamSynchronize(c.copy_in); cudaStreamSynchronize(c.stream); cudaStreamSynchronize(c.copy_out);
    if (result == FPNGB_OK) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) result = 1000 + (int)e; }
    return result;
}

// Per-kernel device times (ms) averaged over the encode calls since the last read; slots in ProfSlot order:
// hist, huffman, scan, offsets, pack, adler, crc.  Returns the number of calls averaged.
FPNGB_API int fpngb_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    g_profile = on != 0;
    return 0;
}
FPNGB_API int fpngb_profile_read(float* ms, int nslots)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    double sum[kProfSlots] = {0};
    uint64_t first_call = ~0ull, last_call = 0;
    for (ProfSet* ps : g_prof_sets) {
        int last = 0;   // index of the last recorded event
        cudaEventSynchronize(ps->ev[0]);
        for (int i = 0; i < kProfSlots; i++) {
            if (!ps->used[i]) continue;
            float t = 0;
            cudaEventSynchronize(ps->ev[i + 1]);
            // the CRC of a chunk may run on the side stream, overlapped with the next chunk: its own start event brackets it
            cudaEvent_t from = (i == kProfCrc && ps->crc_side) ? ps->crc_start : ps->ev[last];
            if (cudaEventElapsedTime(&t, from, ps->ev[i + 1]) == cudaSuccess) sum[i] += t;
            if (!(i == kProfCrc && ps->crc_side)) last = i + 1;
        }
        if (ps->call < first_call) first_call = ps->call;
        if (ps->call > last_call) last_call = ps->call;
        g_prof_free.push_back(ps);
    }
    const int calls = g_prof_sets.empty() ? 0 : (int)(last_call - first_call + 1);
    g_prof_sets.clear();
    for (int i = 0; i < nslots && i < kProfSlots; i++) ms[i] = calls ? (float)(sum[i] / calls) : 0.f;
    return calls;
}

// fpng_crc32 / fpng_adler32 utilities (src/fpng.h:26-31) on host buffers.  Buffers of at least 4 KiB go through the
// device kernels; shorter ones (chunk headers, IHDR) and calls made before fpngb_init() use the host table that the
// container code needs anyway (documented in include/fpng_b200.h).  The *_ex forms report device errors; the plain forms
// keep the reference's signatures and, because 0 is a valid checksum, report a device failure on stderr before returning 0.
int fpngb_crc32_ex(const void* data, size_t size, uint32_t prev, uint32_t* out)
{
    if (!out) return FPNGB_ERR_INVALID_ARG;
    *out = prev;
    if (!data || !size) return FPNGB_OK;
    if (!g_ctx.ready || size < 4096 || size > 0xFFFFFF00ull) { *out = host_crc32(data, size, prev); return FPNGB_OK; }
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    const size_t padded = align_up(size + 16, 16);
    int rc = c.dev_in.reserve(padded + 256); if (rc) return rc;
    ImageState hst{}; hst.zsize = (uint32_t)size - kPngHeaderSize;       // kernel computes L = 58 + zsize = size
    ImageState* dst = (ImageState*)((uint8_t*)c.dev_in.p + padded);
    FPNGB_CUDA_OK(cudaMemcpyAsync(c.dev_in.p, data, size, cudaMemcpyHostToDevice, c.stream));
    FPNGB_CUDA_OK(cudaMemcpyAsync(dst, &hst, sizeof hst, cudaMemcpyHostToDevice, c.stream));
    CrcParams cp{};
    cp.out = (uint8_t*)c.dev_in.p; cp.out_stride = 0; cp.st = dst; cp.max_tiles = crc_ctas_for(size); cp.msg_start = 0; cp.init_xor = ~prev;
    launch_crc(cp, 1, c.stream);
    count_launch(1);
    FPNGB_CUDA_OK(cudaGetLastError());
    uint8_t be[4] = {0, 0, 0, 0};
    FPNGB_CUDA_OK(cudaMemcpyAsync(be, (uint8_t*)c.dev_in.p + size, 4, cudaMemcpyDeviceToHost, c.stream));
    FPNGB_CUDA_OK(cudaStreamSynchronize(c.stream));
    *out = ((uint32_t)be[0] << 24) | ((uint32_t)be[1] << 16) | ((uint32_t)be[2] << 8) | be[3];
    return FPNGB_OK;
}

int fpngb_adler32_ex(const void* data, size_t size, uint32_t adler, uint32_t* out)
{
    if (!out) return FPNGB_ERR_INVALID_ARG;
    *out = adler;
    uint32_t a = adler & 0xFFFF, b = adler >> 16;
    if (!data || !size) return FPNGB_OK;
    if (!g_ctx.ready || size < 4096) {
        const uint8_t* p = (const uint8_t*)data;
        for (size_t i = 0; i < size; i++) { a = (a + p[i]) % kAdlerMod; b = (b + a) % kAdlerMod; }
        *out = (b << 16) | a;
        return FPNGB_OK;
    }
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    const size_t padded = align_up(size, 256), nchunks = (size + adler_chunk_bytes() - 1) / adler_chunk_bytes();
    int rc = c.dev_in.reserve(padded + nchunks * 8); if (rc) return rc;
    uint2* dpart = (uint2*)((uint8_t*)c.dev_in.p + padded);
    std::vector<uint2> part(nchunks);
    FPNGB_CUDA_OK(cudaMemcpyAsync(c.dev_in.p, data, size, cudaMemcpyHostToDevice, c.stream));
    launch_adler_buffer((const uint8_t*)c.dev_in.p, size, dpart, c.stream);
    count_launch(1);
    FPNGB_CUDA_OK(cudaGetLastError());
    FPNGB_CUDA_OK(cudaMemcpyAsync(part.data(), dpart, nchunks * 8, cudaMemcpyDeviceToHost, c.stream));
    FPNGB_CUDA_OK(cudaStreamSynchronize(c.stream));
    unsigned long long A = a, B = b;
    for (size_t i = 0; i < nchunks; i++) {
        const unsigned long long len = (i + 1 < nchunks) ? adler_chunk_bytes() : size - i * adler_chunk_bytes();
        B = (B + (len % kAdlerMod) * A + part[i].y) % kAdlerMod;
        A = (A + part[i].x) % kAdlerMod;
    }
    *out = (uint32_t)((B << 16) | A);
    return FPNGB_OK;
}

uint32_t fpngb_crc32(const void* data, size_t size, uint32_t prev)
{
    uint32_t v = 0;
    const int rc = fpngb_crc32_ex(data, size, prev, &v);
    if (rc) { fprintf(stderr, "fpng_b200: fpng_crc32 failed on the device (error %d); the returned value is NOT a checksum\n", rc); return 0; }
    return v;
}

uint32_t fpngb_adler32(const void* data, size_t size, uint32_t adler)
{
    uint32_t v = 0;
    const int rc = fpngb_adler32_ex(data, size, adler, &v);
    if (rc) { fprintf(stderr, "fpng_b200: fpng_adler32 failed on the device (error %d); the returned value is NOT a checksum\n", rc); return 0; }
    return v;
}

int fpngb_compact_batch_device(const void* d_files, size_t stride, const uint32_t* d_sizes, uint32_t n, void* d_dst, size_t dst_cap,
                               uint64_t* d_offsets, void* stream)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!d_files || !d_sizes || !d_dst || !d_offsets || n == 0 || n > 65535) return FPNGB_ERR_INVALID_ARG;
    if (stride % 16 || (uintptr_t)d_files % 16 || (uintptr_t)d_dst % 16) return FPNGB_ERR_ALIGNMENT;
    FPNGB_CUDA_OK(cudaSetDevice(g_ctx.device));
    launch_compact((const uint8_t*)d_files, stride, d_sizes, n, (uint8_t*)d_dst, dst_cap, (unsigned long long*)d_offsets, (cudaStream_t)stream);
    count_launch(2);
    FPNGB_CUDA_OK(cudaGetLastError());
    return FPNGB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Static-table training (reference: fpng_test -t, src/fpng_test.cpp:766-963, and
// create_dynamic_block_prefix, src/fpng.cpp:909-988).  The GPU part is the 288-bin histogram kernel of 2-pass mode.
// ---------------------------------------------------------------------------------------------------------
// Adds, for every image of a device-resident batch, its 16-bit scaled symbol counts to counts[288] -- exactly what the
// reference accumulates in g_huff_counts while encoding with FPNG_ENCODE_SLOWER (fpng.cpp:751-755 after 1092-1094).
int fpngb_train_accumulate_device(const void* d_pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h, uint32_t chans,
                                  uint64_t* counts, void* stream)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!d_pixels || !counts || n == 0 || !valid_dims(w, h, chans) || image_stride < (size_t)w * h * chans || n > 65535) return FPNGB_ERR_INVALID_ARG;
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    cudaStream_t s = (cudaStream_t)stream;
    Workspace ws;
    int rc = carve_workspace(c, n, h, w, true, ws); if (rc) return rc;
    rc = c.ws_acquire(s); if (rc) return rc;
    ScanParams sp{};
    sp.pixels = (const uint8_t*)d_pixels; sp.image_stride = image_stride; sp.w = w; sp.h = h; sp.books = c.d_static_books; sp.book_stride = 0;
    sp.row_bits = ws.row_bits; sp.row_adler = ws.row_adler; sp.st = ws.st; sp.hist = ws.hist; sp.lane_ofs = ws.lane_ofs; sp.lane_ofs_pitch = ws.lane_ofs_pitch;
    FPNGB_CUDA_OK(cudaMemsetAsync(ws.hist, 0, (size_t)n * 288 * 4, s));
    static const bool force_generic_t = getenv("FPNGB_FORCE_GENERIC") && atoi(getenv("FPNGB_FORCE_GENERIC")) != 0;
    if (!force_generic_t) launch_hist16(sp, n, chans, s);
    else launch_scan(sp, n, chans, pick_load_mode(d_pixels, image_stride, w, chans), true, s);
    count_launch(1);
    std::vector<uint32_t> hh((size_t)n * 288);
    FPNGB_CUDA_OK(cudaMemcpyAsync(hh.data(), ws.hist, hh.size() * 4, cudaMemcpyDeviceToHost, s));
    rc = c.ws_release(s); if (rc) return rc;
    FPNGB_CUDA_OK(cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n; i++) {
        uint32_t* f = &hh[(size_t)i * 288];
        f[256] = 1;                                                   // fpng.cpp:1092
        uint64_t total = 0;
        for (int k = 0; k < 288; k++) total += f[k];
        for (int k = 0; k < 288; k++) if (f[k]) { uint64_t v = (uint64_t)f[k] * 65535u / total; counts[k] += v < 1 ? 1 : v; }   // fpng.cpp:890
    }
    return FPNGB_OK;
}

// counts[288] -> pre-serialised block header ("prefix": zlib header, BFINAL, BTYPE=2, table description), the bits left
// over after the last whole byte, and the code table.  Same signature/semantics as the reference's
// create_dynamic_block_prefix (fpng.cpp:910).  The table construction runs in the Huffman-build kernel.
int fpngb_create_dynamic_block_prefix(const uint64_t* counts, uint32_t chans, uint8_t* prefix, size_t prefix_cap, size_t* prefix_len,
                                      uint64_t* bit_buf, int* bit_buf_size, uint32_t* codes288, uint8_t* sizes288)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!counts || !prefix || !prefix_len || !bit_buf || !bit_buf_size || (chans != 3 && chans != 4)) return FPNGB_ERR_INVALID_ARG;
    uint32_t freq[288];
    for (int i = 0; i < 288; i++) freq[i] = (uint32_t)counts[i];       // fpng.cpp:933 (the reference stores the unshifted count)
    for (int i = 0; i <= 256; i++) if (!freq[i]) freq[i] = 1;          // every literal / EOB codable (fpng.cpp:943-947)
    for (uint32_t len = chans; len <= 258; len += chans) { uint32_t sym, xb, xv; deflate_len_code(len, sym, xb, xv); if (!freq[sym]) freq[sym] = 1; }
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    int rc = c.dev_in.reserve(288 * 4 + sizeof(CodeBook) + 512); if (rc) return rc;
    uint32_t* d_hist = (uint32_t*)c.dev_in.p;
    CodeBook* d_book = (CodeBook*)((uint8_t*)c.dev_in.p + 2048);
    FPNGB_CUDA_OK(cudaMemcpyAsync(d_hist, freq, sizeof freq, cudaMemcpyHostToDevice, c.stream));
    HuffParams hp{d_hist, d_book, chans, 1u};
    launch_huffman_build(hp, 1, c.stream);
    count_launch(1);
    static CodeBook hb;
    FPNGB_CUDA_OK(cudaMemcpyAsync(&hb, d_book, sizeof hb, cudaMemcpyDeviceToHost, c.stream));
    FPNGB_CUDA_OK(cudaStreamSynchronize(c.stream));
    const size_t nbytes = hb.hdr_bits / 8;
    if (prefix_cap < nbytes) return FPNGB_ERR_BUFFER_TOO_SMALL;
    memcpy(prefix, hb.hdr, nbytes);
    *prefix_len = nbytes;
    *bit_buf_size = (int)(hb.hdr_bits & 7);
    *bit_buf = hb.hdr[nbytes] & ((1u << (hb.hdr_bits & 7)) - 1u);
    // the exported code table covers every symbol with a code (canonical assignment, fpng.cpp:701-708), including
    // length symbols this encoder never emits
    uint16_t codes16[288];
    canonical_codes(hb.sym_size, 288, codes16);
    for (int i = 0; i < 288; i++) {
        if (sizes288) sizes288[i] = hb.sym_size[i];
        if (codes288) codes288[i] = codes16[i];
    }
    return FPNGB_OK;
}

// Install a trained 1-pass table (prefix bytes + leftover bits, as produced above or by the reference's trainer) for
// `chans`; nbytes == 0 restores the built-in table.  Files written with it are ordinary fpng files.
int fpngb_set_static_table(uint32_t chans, const uint8_t* prefix, size_t nbytes, uint32_t bit_buf, uint32_t bit_buf_size)
{
    if (!g_ctx.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (chans != 3 && chans != 4) return FPNGB_ERR_INVALID_ARG;
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    static CodeBook cb;                                                // static: still alive while the async copy runs
    bool ok;
    if (nbytes == 0) ok = chans == 3 ? build_static_book(cb, kStaticHdrRGB, sizeof kStaticHdrRGB, kStaticHdrRGBTail, kStaticHdrRGBTailBits, 3)
                                     : build_static_book(cb, kStaticHdrRGBA, sizeof kStaticHdrRGBA, kStaticHdrRGBATail, kStaticHdrRGBATailBits, 4);
    else {
        if (!prefix || nbytes + 1 > kMaxHdrBytes || bit_buf_size > 7) return FPNGB_ERR_INVALID_ARG;
        ok = build_static_book(cb, prefix, (uint32_t)nbytes, bit_buf, bit_buf_size, chans);
    }
    if (!ok) return FPNGB_ERR_INVALID_ARG;
    FPNGB_CUDA_OK(cudaDeviceSynchronize());                            // no encode may be using the old table
    c.h_static_books[chans == 4 ? 1 : 0] = cb;
    FPNGB_CUDA_OK(cudaMemcpy(c.d_static_books + (chans == 4 ? 1 : 0), &cb, sizeof cb, cudaMemcpyHostToDevice));
    return FPNGB_OK;
}

// Binds the calling thread (and the threads it creates later) to the CPUs of the NUMA node the library's GPU hangs off, so that
// pinned staging buffers allocated afterwards (first touch) and the copy-issuing thread sit next to the GPU's PCIe root port.
// With 8 ranks on a two-socket host this is what keeps H2D/D2H of the *_host entry points from crossing the socket
// interconnect.  Returns the NUMA node (>= 0), or -1 when the topology is not exposed (single node, container without
// sysfs): then nothing is changed.
int fpngb_bind_host_thread_to_device_numa(void)
{
    if (!g_ctx.ready) return -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, g_ctx.device) != cudaSuccess) return -1;
    for (char* q = bus; *q; q++) if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) return -1;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return -1;
    char list[4096] = {0};
    const size_t got = fread(list, 1, sizeof list - 1, f);
    fclose(f);
    if (!got) return -1;
    cpu_set_t set; CPU_ZERO(&set);
    cpu_set_t allowed; CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return -1;
    int count = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { } else if (sscanf(tok, "%d", &a) == 1) b = a; else continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &set); count++; }
    }
    if (!count) return -1;                                             // the node's CPUs are outside this process's cpuset: leave it alone
    if (sched_setaffinity(0, sizeof set, &set) != 0) return -1;
    return node;
}

void* fpngb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void fpngb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
