// fpng_b200/csrc/decode_kernels.cu -- sm_100a kernels of the fpng decode hot path (fpng-written files only, like the
// reference: src/fpng.cpp:2209-2901).
//
// The reference decodes one serial Huffman bit-string (no row index exists in the format).  Here:
//   D1 decode_prepare_kernel   one warp per file: zlib/Deflate block header checks, code-length parse, fpng table
//                              constraints (fpng.cpp:1954-2076), 4096-entry literal/length LUT (fpng.cpp:1836-1895)
//   D2 decode_tokens_kernel    one CTA per file walks the stream in chunks of kDecThreads subsequences of kSubBits bits;
//                              inside a chunk every thread decodes its subsequence speculatively and the chunk iterates
//                              "restart at the predecessor's exit point" until no exit moves (self-synchronising prefix
//                              code; the first thread of a chunk always starts exactly), then a block scan of output byte
//                              counts places every subsequence in the filtered stream and a last decode writes the delta
//                              bytes (literals) / replicates the previous delta pixel (RLE matches, fpng.cpp:2289-2388)
//   D3 unfilter_kernel         inverse PNG filter 2 = running sum down each byte column (fpng.cpp:2439-2466 fuses it into
//                              the serial loop) + 24<->32bpp conversion (alpha 0xFF / dropped)
//   D2' decode_stored_kernel   stored-block files (fpng.cpp:2107-2207)
// Every stream violation the reference rejects ends in status FPNG_DECODE_NOT_FPNG (fpng.cpp:3131-3136).
#include "row_walk.cuh"
#include "kernels.cuh"
#include "decode.cuh"

namespace fpngb {

__constant__ uint16_t c_len_base[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
__constant__ uint8_t c_len_xbits[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};

// ------------------------------------------------------------------------------------------------
// bit reader over the zlib bytes of one file; reads past `limit` bytes return zero bits
// ------------------------------------------------------------------------------------------------
struct BitSrc {
    const uint8_t* z;          // zlib stream start (file + idat_ofs + 8), any alignment
    uint32_t limit;            // readable bytes from z (to the end of the file)
    __device__ __forceinline__ uint32_t word_at(uint32_t byte_ofs) const
    {
        // unaligned little-endian 32-bit read assembled from aligned words (z itself may be unaligned)
        const uintptr_t a = (uintptr_t)(z + byte_ofs);
        const uint32_t* p = (const uint32_t*)(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        uint32_t lo = 0, hi = 0;
        if (byte_ofs < limit) lo = __ldg(p);
        if (sh && byte_ofs + 4 - (sh >> 3) < limit) hi = __ldg(p + 1);
        uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
        if (byte_ofs + 4 > limit) { const uint32_t valid = byte_ofs < limit ? limit - byte_ofs : 0; v &= valid >= 4 ? 0xFFFFFFFFu : ((1u << (8 * valid)) - 1u); }
        return v;
    }
};

struct BitCursor {
    unsigned long long buf; uint32_t cnt; uint32_t next_byte;   // next_byte: offset of the next unread byte
    unsigned long long pos;                                     // absolute bit position of buf's bit 0
    __device__ __forceinline__ void seek(const BitSrc& s, unsigned long long bitpos)
    {
        pos = bitpos;
        const uint32_t b = (uint32_t)(bitpos >> 3), sh = (uint32_t)(bitpos & 7);
        buf = ((unsigned long long)s.word_at(b) | ((unsigned long long)s.word_at(b + 4) << 32)) >> sh;
        cnt = 64 - sh; next_byte = b + 8;
    }
    __device__ __forceinline__ void refill(const BitSrc& s)
    {
        if (cnt <= 32) { buf |= (unsigned long long)s.word_at(next_byte) << cnt; cnt += 32; next_byte += 4; }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void skip(uint32_t n) { buf >>= n; cnt -= n; pos += n; }
    __device__ __forceinline__ uint32_t get(const BitSrc& s, uint32_t n) { refill(s); const uint32_t v = peek(n); skip(n); return v; }
};

// ------------------------------------------------------------------------------------------------
// D1: per-file block header -> LUT
// ------------------------------------------------------------------------------------------------
// Canonical-code LUT with the reference's acceptance rule (fpng.cpp:1836-1895): complete code, or exactly one code.
// Runs on one warp; `sizes` in shared memory.
__device__ static bool build_lut_warp(const uint8_t* sizes, uint32_t nsyms, uint16_t* lut, uint32_t lut_bits, uint32_t lane, uint32_t* s_next /*17*/)
{
    __syncwarp();
    if (lane == 0) {
        uint32_t cnt[16];
        for (int i = 0; i < 16; i++) cnt[i] = 0;
        for (uint32_t i = 0; i < nsyms; i++) cnt[sizes[i]]++;
        uint32_t total = 0;
        s_next[0] = s_next[1] = 0;
        for (int l = 1; l <= 15; l++) { total = (total + cnt[l]) << 1; s_next[l + 1] = total; }
        uint32_t ok = 1;
        if (total != 0x10000u) {
            uint32_t used = 0;
            for (int l = 15; l >= 1; l--) used += cnt[l];
            ok = used == 1;
        }
        s_next[0] = ok;
    }
    __syncwarp();
    if (!s_next[0]) return false;
    const uint32_t lut_size = 1u << lut_bits;
    for (uint32_t i = lane; i < lut_size; i += 32) lut[i] = 0;
    __syncwarp();
    // canonical code of symbol i = first_code[len] + (number of earlier symbols of the same length); symbols are spread
    // over the lanes, each lane ranks and replicates its own
    for (uint32_t i = lane; i < nsyms; i += 32) {
        const uint32_t l = sizes[i];
        if (!l || l > lut_bits) continue;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < i; j++) rank += sizes[j] == l;
        const uint32_t code = __brev(s_next[l] + rank) >> (32 - l);
        for (uint32_t c = code; c < lut_size; c += 1u << l) lut[c] = (uint16_t)(i | (l << 9));
    }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(32) decode_prepare_kernel(DecodeParams p)
{
    __shared__ uint8_t s_sizes[288 + 32];
    __shared__ uint8_t s_cl[19];
    __shared__ uint16_t s_cllut[128];
    __shared__ uint32_t s_next[17];
    __shared__ uint32_t s_fail;
    __shared__ uint16_t s_base[4096];
    const uint32_t f = blockIdx.x, lane = threadIdx.x;
    const FileDesc fd = p.files[f];
    DecodeState* st = p.state + f;
    const uint8_t* z = p.d_files + (size_t)f * p.file_stride + fd.idat_ofs + 8;
    BitSrc src{z, fd.file_size - (fd.idat_ofs + 8)};
    uint32_t* lut = p.luts + (size_t)f * 4096;

    if (lane == 0) {
        s_fail = 0;
        st->status = 0; st->stored = 0; st->token_start = 0; st->out_bytes = 0; st->end_byte = 0;
        // zlib header and block type (fpng.cpp:2219-2244)
        if (fd.idat_len < 7 || z[0] != 0x78 || z[1] != 0x01) s_fail = 1;
        else if ((z[2] & 6) == 0) st->stored = 1;
        else if ((z[2] & 7) != 5) s_fail = 1;           // BFINAL = 1, BTYPE = 2
    }
    __syncwarp();
    if (s_fail) { if (lane == 0) st->status = 1; return; }
    if (st->stored) return;

    // HLIT / HDIST / HCLEN and the code-length code
    BitCursor bc; bc.seek(src, 16 + 3);
    const uint32_t nlit = bc.get(src, 5) + 257, ndist = bc.get(src, 5) + 1, nclen = bc.get(src, 4) + 4;
    if (lane < 19) s_cl[lane] = 0;
    __syncwarp();
    if (lane == 0) {
        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (uint32_t i = 0; i < nclen; i++) s_cl[order[i]] = (uint8_t)bc.get(src, 3);
    }
    __syncwarp();
    bool ok = nlit + ndist <= 288 + 32;
    ok = ok && build_lut_warp(s_cl, 19, s_cllut, 7, lane, s_next);
    if (!ok) { if (lane == 0) st->status = 1; return; }

    if (lane == 0) {
        for (uint32_t i = 0; i < 288 + 32; i++) s_sizes[i] = 0;
        uint32_t cur = 0, fail = 0;
        const uint32_t total = nlit + ndist;
        while (cur < total && !fail) {
            bc.refill(src);
            const uint32_t e = s_cllut[bc.peek(7)], l = e >> 9, s = e & 511;
            if (!l) { fail = 1; break; }
            bc.skip(l);
            if (s <= 15) { if (s > 12) { fail = 1; break; } s_sizes[cur++] = (uint8_t)s; continue; }   // fpng.cpp:2007
            uint32_t rep, val = 0;
            if (s == 16) { rep = bc.get(src, 2) + 3; if (!cur) { fail = 1; break; } val = s_sizes[cur - 1]; }
            else if (s == 17) rep = bc.get(src, 3) + 3;
            else rep = bc.get(src, 7) + 11;
            if (cur + rep > total) { fail = 1; break; }
            while (rep--) s_sizes[cur++] = (uint8_t)val;
        }
        // distance code constraints (fpng.cpp:2058-2074)
        uint32_t ones = 0;
        for (uint32_t i = 0; i < ndist && !fail; i++) ones += s_sizes[nlit + i] == 1;
        if (!fail) {
            if (ones < 1 || ones > 2) fail = 1;
            else if (s_sizes[nlit + p.chans - 1] != 1) fail = 1;
            else if (ones == 2 && s_sizes[nlit + p.chans] != 1) fail = 1;
        }
        for (uint32_t i = nlit; i < 288; i++) s_sizes[i] = 0;
        s_fail = fail;
        st->token_start = bc.pos;
    }
    __syncwarp();
    if (s_fail) { if (lane == 0) st->status = 1; return; }
    if (!build_lut_warp(s_sizes, 288, s_base, 12, lane, s_next)) { if (lane == 0) st->status = 1; return; }
    // Fuse a second literal into the entry when both codes fit in the 12 index bits (the reference augments its table the
    // same way, fpng.cpp:2079-2102):  sym0 | len0 << 9 | sym1 << 13 | len1 << 21.
    for (uint32_t i = lane; i < 4096; i += 32) {
        const uint32_t e0 = s_base[i], sym0 = e0 & 511u, len0 = e0 >> 9;
        uint32_t e = sym0 | (len0 << 9);
        if (len0 && sym0 < 256u) {
            const uint32_t e1 = s_base[i >> len0], sym1 = e1 & 511u, len1 = e1 >> 9;
            if (len1 && sym1 < 256u && len0 + len1 <= 12u) e |= (sym1 << 13) | (len1 << 21);
        }
        lut[i] = e;
    }
    // Fast table of the scan / write loops (one shared-memory look-up per 12 stream bits, one uniform code path):
    //   bits 0..3  L     stream bits consumed by the whole entry (0: not fast -- end of block, invalid code, a match whose
    //                    code + extra bits + distance bit do not fit in the 12 index bits; the loops fall back to `lut`)
    //   bits 4..5  cnt   number of literals (1..3), 0 = one match
    //   bits 8..31       the literals, first one in the low byte / the match's run length in bytes
    // followed by the 256 literal code sizes (the loops split a multi-literal entry at a subsequence boundary with them).
    uint32_t* fast = p.fast + (size_t)f * kFastWords;
    for (uint32_t i = lane; i < 4096; i += 32) {
        const uint32_t e0 = s_base[i], sym0 = e0 & 511u, len0 = e0 >> 9;
        uint32_t fe = 0;
        if (len0 && sym0 < 256u) {
            uint32_t L = len0, cnt = 1, P = sym0;
            for (uint32_t k = 1; k < 3; k++) {
                const uint32_t e1 = s_base[i >> L], sym1 = e1 & 511u, len1 = e1 >> 9;
                if (!len1 || sym1 >= 256u || L + len1 > 12u) break;
                P |= sym1 << (8u * k); L += len1; cnt++;
            }
            fe = L | (cnt << 4) | (P << 8);
        } else if (len0 && sym0 > 256u && sym0 <= 285u) {
            const uint32_t xb = c_len_xbits[sym0 - 257u];
            if (len0 + xb + 1u <= 12u) fe = (len0 + xb + 1u) | ((c_len_base[sym0 - 257u] + ((i >> len0) & ((1u << xb) - 1u))) << 8);
        }
        fast[i] = fe;
    }
    for (uint32_t i = lane; i < 64; i += 32)
        fast[4096 + i] = s_sizes[4 * i] | (s_sizes[4 * i + 1] << 8) | (s_sizes[4 * i + 2] << 16) | ((uint32_t)s_sizes[4 * i + 3] << 24);
}

// ------------------------------------------------------------------------------------------------
// D2: token decode.  The token bit-string of a file is cut into subsequences of kSubBits bits (absolute multiples of
// kSubBits in the aligned-stream bit coordinate).  Three kernels:
//   D2a decode_scan_kernel   one thread per subsequence, all files and subsequences in parallel: the thread starts
//                            kPreRoll bits early (speculatively), lets the self-synchronising prefix code lock on, takes
//                            the first token boundary inside its subsequence as its start, then decodes to its exit (first
//                            boundary in the next subsequence) counting output bytes and remembering the last 4 literals.
//   D2b decode_link_kernel   one CTA per file walks the subsequences in order, 1024 at a time: verifies that every
//                            subsequence starts where its predecessor exits (the first one starts exactly, so a verified
//                            chain is the true parse), repairs the rare mismatches by re-decoding, and scans the byte
//                            counts / literal windows into the output offset and "previous pixel" of every subsequence.
//   D2c decode_write_kernel  one thread per live subsequence decodes again and writes the delta bytes.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRelEnd = 0xFFFFFFFFu;              // exit flag (relative positions): saw the end-of-block code
constexpr uint32_t kRelErr = 0xFFFFFFFEu;              // exit flag: invalid code
constexpr unsigned long long kPosEnd = ~0ull;          // same flags in absolute (64-bit) form
constexpr unsigned long long kPosErr = ~0ull - 1;

// Aligned view of one file's zlib stream: 32-bit words starting at the 4-byte boundary at or before the stream.
struct Stream {
    const uint32_t* words;     // aligned base
    uint32_t max_widx;         // last readable word index (reads beyond are clamped: garbage but in bounds)
    uint32_t bit0;             // aligned-stream bit position of zlib bit 0 (0, 8, 16 or 24)
    // optional shared-memory window over words [sw_base, sw_base + sw_count): the scan and write kernels stage the 32 KiB of
    // stream their 256 subsequences walk (coalesced loads, one pad word per 32 so that threads 32 words apart hit different
    // banks); a refill is then a ~30-cycle shared load instead of an L1/L2 round trip in the middle of a dependent chain
    const uint32_t* swin; uint32_t sw_base, sw_count;
    __device__ __forceinline__ uint32_t word(uint32_t widx) const
    {
        const uint32_t si = widx - sw_base;
        if (si < sw_count) return swin[si + (si >> 5)];
        return __ldg(words + min(widx, max_widx));
    }
};
constexpr uint32_t kWinLead = 8, kWinTail = 8;                                   // words before / after the CTA's 256 subsequences
constexpr uint32_t kWinWords = kWinLead + kDecThreads * (kSubBits / 32) + kWinTail;
constexpr uint32_t kWinSmemWords = kWinWords + kWinWords / 32 + 1;

__device__ __forceinline__ void cp_async4(uint32_t* smem_dst, const uint32_t* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16g(uint32_t* smem_dst, const uint32_t* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// cooperative load of the CTA's stream window; `first_sub` = index of the CTA's first subsequence
__device__ __forceinline__ void stage_window(Stream& st, uint32_t* s_win, unsigned long long first_sub)
{
    const unsigned long long w0 = first_sub * (kSubBits / 32);
    const uint32_t base = (uint32_t)(w0 >= kWinLead ? w0 - kWinLead : 0ull);
    // word i lives at s_win[i + i / 32]; the pad slot in front of every 32-word block holds a COPY of the block's first word, so
    // that (address, address + 1) always are two consecutive stream words (win_peek below reads 32 bits at any bit position)
    // asynchronous copies (LDGSTS): all of a thread's ~33 words are in flight at once instead of one load -> store round trip each
    // (ncu: a quarter of the scan kernel's stall samples sat on the staging stores); the caller waits with stage_wait()
    for (uint32_t i = threadIdx.x; i < kWinWords; i += blockDim.x) {
        const uint32_t* src = st.words + min(base + i, st.max_widx);
        cp_async4(s_win + i + (i >> 5), src);
        if ((i & 31u) == 0u && i) cp_async4(s_win + i + (i >> 5) - 1u, src);
    }
    st.swin = s_win; st.sw_base = base; st.sw_count = kWinWords;
}

// the file's fast table (kFastWords words, 16-byte aligned) into shared memory, asynchronously
__device__ __forceinline__ void stage_fast_table(uint32_t* s_fast, const uint32_t* __restrict__ g_fast)
{
    for (uint32_t i = threadIdx.x; i < kFastWords / 4u; i += blockDim.x) cp_async16g(s_fast + 4u * i, g_fast + 4u * i);
}
__device__ __forceinline__ void stage_wait()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
}

__device__ __forceinline__ Stream open_stream(const uint8_t* file, const FileDesc& fd)
{
    Stream st;
    const uint32_t zofs = fd.idat_ofs + 8;
    st.words = reinterpret_cast<const uint32_t*>(file + (zofs & ~3u));
    st.bit0 = (zofs & 3u) * 8u;
    st.max_widx = ((fd.file_size + 3u) >> 2) - 1u - (zofs >> 2);     // file buffers are padded to a multiple of 4
    st.swin = nullptr; st.sw_base = 0; st.sw_count = 0;
    return st;
}

struct Cursor {
    unsigned long long buf; uint32_t cnt, widx;
    uint32_t ahead;            // the next stream word, loaded one refill early so that its latency overlaps the decoding of ~32 bits
    __device__ __forceinline__ void seek(const Stream& st, unsigned long long abs_bit)
    {
        widx = (uint32_t)(abs_bit >> 5);
        const uint32_t sh = (uint32_t)abs_bit & 31u;
        const uint32_t lo = st.word(widx), hi = st.word(widx + 1u);
        buf = (((unsigned long long)hi << 32) | lo) >> sh;
        cnt = 64u - sh; widx += 2u;
        ahead = st.word(widx);
    }
    __device__ __forceinline__ void refill(const Stream& st)
    {
        if (cnt <= 32u) {
            buf |= (unsigned long long)ahead << cnt; cnt += 32u; widx++;
            ahead = st.word(widx);
        }
    }
    __device__ __forceinline__ void skip(uint32_t n) { buf >>= n; cnt -= n; }
};

struct SubScan {
    uint32_t first;      // rel. position of the first token boundary >= lo (the subsequence's start)
    uint32_t exit;       // rel. position of the first token boundary >= hi, or kRelEnd / kRelErr
    uint32_t eob_end;    // rel. position right after the EOB code (exit == kRelEnd)
    uint32_t n_out;      // bytes produced by tokens starting in [first, hi)
    uint32_t nlit;       // literal bytes among them (saturating at 255)
    uint32_t lits;       // last <= 4 literal bytes, most recent in the top byte
};

// Fused LUT entry: sym0 (9 bits) | len0 << 9 (4 bits) | sym1 << 13 (8 bits) | len1 << 21 (4 bits; 0 = no second literal)
__device__ __forceinline__ uint32_t lut_sym0(uint32_t e) { return e & 511u; }
__device__ __forceinline__ uint32_t lut_len0(uint32_t e) { return (e >> 9) & 15u; }
__device__ __forceinline__ uint32_t lut_sym1(uint32_t e) { return (e >> 13) & 255u; }
__device__ __forceinline__ uint32_t lut_len1(uint32_t e) { return e >> 21; }

// Write side of the token decoder: consecutive delta bytes of a row are combined into aligned 32-bit stores (the words
// a subsequence shares with its neighbours at either end, and row tails, go out as single bytes).
struct DeltaSink {
    uint8_t* delta; uint32_t pitch, bpl, h;
    uint32_t row, col;            // position in the filtered stream; col 0 is the filter byte
    uint8_t* ptr;                 // address of the next data byte of the current row
    uint32_t acc;                 // bytes [word(ptr), ptr) of the aligned word being filled, already in place
    uint32_t first_off;           // offset inside that word of the first byte this thread owns (0 unless it started mid-word)
    __device__ __forceinline__ void init(uint8_t* d, uint32_t pitch_, uint32_t bpl_, uint32_t h_, unsigned long long out_pos)
    {
        delta = d; pitch = pitch_; bpl = bpl_; h = h_;
        row = (uint32_t)(out_pos / (bpl + 1ull)); col = (uint32_t)(out_pos % (bpl + 1ull));
        ptr = delta + (size_t)row * pitch + (col ? col - 1u : 0u);
        acc = 0; first_off = (uint32_t)((uintptr_t)ptr & 3u);
    }
    // store the bytes gathered so far for the (incomplete or partly owned) current word, one by one
    __device__ __forceinline__ void flush()
    {
        const uint32_t n = (uint32_t)((uintptr_t)ptr & 3u);
        uint8_t* w = ptr - n;
        for (uint32_t i = first_off; i < n; i++) w[i] = (uint8_t)(acc >> (8u * i));
        acc = 0; first_off = n;
    }
    __device__ __forceinline__ void next_row() { flush(); col = 0; row++; ptr = delta + (size_t)row * pitch; first_off = 0; }
    // one data byte (col >= 1, row < h)
    __device__ __forceinline__ void put(uint32_t v)
    {
        acc |= v << (8u * ((uint32_t)(uintptr_t)ptr & 3u));
        ptr++;
        if (((uintptr_t)ptr & 3u) == 0) {                  // word complete
            if (first_off == 0) *reinterpret_cast<uint32_t*>(ptr - 4) = acc;
            else { for (uint32_t i = first_off; i < 4u; i++) ptr[(int)i - 4] = (uint8_t)(acc >> (8u * i)); first_off = 0; }
            acc = 0;
        }
        if (++col > bpl) next_row();
    }
};

// Decode from rel. position `rel` (origin = abs_origin bits in the aligned stream).  Tokens starting before `lo` are
// pre-roll (not counted); tokens starting in [lo, hi) are counted / written.  kWrite writes delta bytes.
template <bool kWrite>
__device__ __forceinline__ SubScan decode_range(const Stream& st, const uint32_t* __restrict__ s_lut, unsigned long long abs_origin,
                                                uint32_t rel, uint32_t lo, uint32_t hi, uint32_t chans,
                                                uint8_t* __restrict__ delta, uint32_t pitch, uint32_t bpl, uint32_t h,
                                                unsigned long long out_pos, uint32_t tail, uint32_t* err)
{
    SubScan r; r.first = lo; r.exit = 0; r.eob_end = 0; r.n_out = 0; r.nlit = 0; r.lits = 0;
    Cursor c; c.seek(st, abs_origin + rel);
    // ---- pre-roll: lock on to the token grid (speculative; an end-of-block or invalid code here means "not locked")
    while (rel < lo) {
        c.refill(st);
        const uint32_t e = s_lut[(uint32_t)c.buf & 4095u];
        uint32_t l = lut_len0(e);
        const uint32_t s = lut_sym0(e);
        if (!l || s == 256u || s > 285u) { rel = lo; c.seek(st, abs_origin + rel); break; }
        if (s > 256u) { const uint32_t xb = c_len_xbits[s - 257u]; l += xb + 1u; }
        c.skip(l); rel += l;
    }
    r.first = rel;
    uint32_t lits = tail;
    DeltaSink sink;
    if (kWrite) sink.init(delta, pitch, bpl, h, out_pos);
    uint32_t n_out = 0, nlit = 0;
    while (rel < hi) {
        c.refill(st);
        const uint32_t e = s_lut[(uint32_t)c.buf & 4095u];
        const uint32_t l0 = lut_len0(e), s = lut_sym0(e);
        if (!l0) { r.exit = kRelErr; r.n_out = n_out; r.nlit = nlit; r.lits = lits; return r; }
        if (s < 256u) {
            // one literal, or two when the fused second literal also starts inside the subsequence
            const uint32_t l1 = lut_len1(e);
            const bool two = l1 && (rel + l0 < hi);
            const uint32_t l = two ? l0 + l1 : l0;
            c.skip(l); rel += l;
            const uint32_t cntl = two ? 2u : 1u;
            n_out += cntl; nlit += cntl;
            if (!kWrite) {
                lits = (lits >> 8) | (s << 24);
                if (two) lits = (lits >> 8) | (lut_sym1(e) << 24);
            } else {
                // first literal, then (predicated) the fused second one; a literal at column 0 is the row's filter byte
                lits = (lits >> 8) | (s << 24);
                if (sink.col == 0) { if (sink.row >= h || s != (sink.row ? 2u : 0u)) { *err = 1; break; } sink.col = 1; }   // fpng.cpp:2264, 2642
                else sink.put(s);
                if (two) {
                    const uint32_t v = lut_sym1(e);
                    lits = (lits >> 8) | (v << 24);
                    if (sink.col == 0) { if (sink.row >= h || v != (sink.row ? 2u : 0u)) { *err = 1; break; } sink.col = 1; }
                    else sink.put(v);
                }
            }
        } else if (s == 256u) {
            c.skip(l0); rel += l0;
            if (kWrite) sink.flush();
            r.exit = kRelEnd; r.eob_end = rel; r.n_out = n_out; r.nlit = nlit; r.lits = lits;
            return r;
        } else {
            if (s > 285u) { r.exit = kRelErr; r.n_out = n_out; r.nlit = nlit; r.lits = lits; return r; }
            c.skip(l0);
            const uint32_t xb = c_len_xbits[s - 257u];
            const uint32_t run = c_len_base[s - 257u] + ((uint32_t)c.buf & ((1u << xb) - 1u));
            c.skip(xb + 1u);                                                       // extra bits + the 1-bit distance code (fpng.cpp:2300)
            rel += l0 + xb + 1u;
            n_out += run;
            if (kWrite) {
                // run of the previous delta pixel: starts on a pixel boundary after at least one pixel of the row, is a
                // whole number of pixels and stays inside the row (fpng.cpp:2302-2315, 2681-2691, 2727)
                const uint32_t row = sink.row, col = sink.col;
                const bool bad = row >= h || col < 1u + chans || ((col - 1u) % chans) != 0 || (run % chans) != 0 || (col - 1u) + run > bpl;
                if (bad) { *err = 1; return r; }
                const uint32_t px = chans == 4 ? lits : (lits >> 8);              // last `chans` literals, oldest in the low byte
                if (chans == 4) {
                    // RGBA pixels are word aligned in the delta rows: one 32-bit store per pixel
                    sink.flush();
                    uint32_t* d = reinterpret_cast<uint32_t*>(sink.ptr);
                    const uint32_t npx = run >> 2;
                    for (uint32_t i = 0; i < npx; i++) d[i] = px;
                    sink.ptr += run; sink.col += run;
                    if (sink.col > bpl) sink.next_row();
                } else {
                    for (uint32_t i = 0; i < run; i += 3) { sink.put(px & 0xFFu); sink.put((px >> 8) & 0xFFu); sink.put(px >> 16); }
                }
            }
        }
    }
    if (kWrite) sink.flush();
    r.exit = rel; r.n_out = n_out; r.nlit = nlit; r.lits = lits;
    return r;
}

// Write pass, lean form: decodes the tokens of one subsequence from its verified start and stores the delta bytes.
// Output state is three 32-bit values (row, data column, pending-byte accumulator) and a row pointer; bytes are gathered into
// aligned 32-bit stores (a thread's unaligned head bytes and a row's tail bytes go out singly, so neighbouring threads
// never touch the same word).  Same acceptance rules as decode_range<true> (fpng.cpp:2264, 2302-2315, 2642, 2681-2691, 2727).
// ---- lean token loops on the shared-memory window -------------------------------------------------------------------
// With the stream staged in shared memory a decoder needs no bit buffer: 32 fresh bits at ANY bit position are two shared
// loads and one funnel shift, and the only state carried from token to token is the bit position itself.
__device__ __forceinline__ uint32_t win_peek(const uint32_t* __restrict__ win, uint32_t q)     // q: bit offset from the window's first word
{
    const uint32_t si = q >> 5, a = si + (si >> 5);
    return __funnelshift_r(win[a], win[a + 1u], q);                         // shift amount taken modulo 32
}

// One step of the fast loops: looks up the entry at bit position q of the window.  A multi-literal entry whose later
// literals could start at or beyond `limit` (the subsequence boundary) is cut down to its first literal.  Returns false when
// the entry is not fast (L = 0); then `w` still holds the 32 stream bits for the fall-back.
struct FastTok { uint32_t L, cnt, payload, w; };
__device__ __forceinline__ bool fast_tok(const uint32_t* __restrict__ win, const uint32_t* __restrict__ s_fast, uint32_t sizes_s, uint32_t q, uint32_t rel, uint32_t limit, FastTok& t)
{
    t.w = win_peek(win, q);
    const uint32_t e = s_fast[t.w & 4095u];
    t.L = e & 15u; t.cnt = (e >> 4) & 3u; t.payload = e >> 8;
    if (t.cnt >= 2u && rel + 12u > limit) {
        t.payload &= 0xFFu; t.cnt = 1u;
        asm("ld.shared.u8 %0, [%1];\n" : "=r"(t.L) : "r"(sizes_s + t.payload));      // code size of the first literal
    }
    return t.L != 0u;
}

// Fall-back of the fast loops: one token through the single-token table in global memory.  Returns the token's length in bits
// (0 = invalid code), sets sym (256 = end of block) and, for a match, run.
__device__ __forceinline__ uint32_t slow_tok(const uint32_t* __restrict__ lutg, uint32_t w, uint32_t& sym, uint32_t& run)
{
    const uint32_t e = __ldg(lutg + (w & 4095u));
    const uint32_t l0 = lut_len0(e);
    sym = lut_sym0(e); run = 0;
    if (!l0 || sym > 285u) return 0u;
    if (sym <= 256u) return l0;
    const uint32_t xb = c_len_xbits[sym - 257u];
    run = c_len_base[sym - 257u] + ((w >> l0) & ((1u << xb) - 1u));
    return l0 + xb + 1u;                                                     // extra bits + the 1-bit distance code (fpng.cpp:2300)
}

// Scan pass on the window: same contract as decode_range<false>.  `O` = bit offset of abs_origin inside the window.
__device__ __forceinline__ SubScan scan_window(const uint32_t* __restrict__ win, uint32_t O, const uint32_t* __restrict__ s_fast,
                                               const uint32_t* __restrict__ lutg, uint32_t rel, uint32_t lo, uint32_t hi)
{
    SubScan r; r.first = lo; r.exit = 0; r.eob_end = 0; r.n_out = 0; r.nlit = 0; r.lits = 0;
    FastTok t;
    const uint32_t sizes_s = (uint32_t)__cvta_generic_to_shared(s_fast + 4096);
    while (rel < lo) {                                                       // pre-roll: lock on to the token grid
        if (fast_tok(win, s_fast, sizes_s, O + rel, rel, lo, t)) { rel += t.L; continue; }
        uint32_t sym, run;
        const uint32_t l = slow_tok(lutg, t.w, sym, run);
        if (!l || sym == 256u) { rel = lo; break; }                          // not locked on
        rel += l;
    }
    r.first = rel;
    uint32_t lits = 0, n_out = 0, nlit = 0;
    while (rel < hi) {
        if (fast_tok(win, s_fast, sizes_s, O + rel, rel, hi, t)) {
            rel += t.L;
            n_out += t.cnt ? t.cnt : t.payload;
            nlit += t.cnt;
            lits = __funnelshift_r(lits, t.payload, 8u * t.cnt);              // the new literals enter at the top (cnt = 0: unchanged)
            continue;
        }
        uint32_t sym, run;
        const uint32_t l = slow_tok(lutg, t.w, sym, run);
        if (!l) { r.exit = kRelErr; r.n_out = n_out; r.nlit = nlit; r.lits = lits; return r; }
        rel += l;
        if (sym == 256u) { r.exit = kRelEnd; r.eob_end = rel; r.n_out = n_out; r.nlit = nlit; r.lits = lits; return r; }
        if (sym < 256u) { n_out++; nlit++; lits = (lits >> 8) | (sym << 24); }   // (cannot happen: every literal is a fast entry)
        else n_out += run;
    }
    r.exit = rel; r.n_out = n_out; r.nlit = nlit; r.lits = lits;
    return r;
}

template <int CHANS>
__device__ __forceinline__ void decode_write_range(const Stream& st, const uint32_t* __restrict__ s_fast, const uint32_t* __restrict__ lutg, unsigned long long abs_origin, uint32_t hi,
                                                   uint8_t* __restrict__ delta, uint32_t pitch, uint32_t bpl, uint32_t h,
                                                   unsigned long long out_pos, uint32_t tail, uint32_t* err)
{
    // the subsequence lies inside the CTA's shared-memory window (stage_window): read it by bit position
    const uint32_t* __restrict__ win = st.swin;
    const uint32_t O = (uint32_t)(abs_origin - 32ull * st.sw_base);
    uint32_t rel = 0, lits = tail;
    uint32_t row = (uint32_t)(out_pos / (bpl + 1ull));
    const uint32_t col0 = (uint32_t)(out_pos % (bpl + 1ull));
    bool need_filter = col0 == 0u;                     // the next byte of the filtered stream is the row's filter byte
    uint32_t dcol = col0 ? col0 - 1u : 0u;             // data bytes of the current row already produced
    uint8_t* rowp = delta + (size_t)row * pitch;
    unsigned long long acc = 0; uint32_t nacc = 0;     // pending bytes: data columns [dcol - nacc, dcol), first one 4-byte aligned
    // Completed words of the literal fast path wait in a 4-deep register queue (newest in q3) and leave as ONE 128-bit store once
    // they fill a 16-byte aligned group: the write pass is bound by store transactions (every lane writes its own region, so a
    // 32-bit store costs a whole L1 tag cycle per lane), not by decoding.  Every other path flushes the queue first.
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0, nq = 0;   // queued words: data columns [dcol - nacc - 4 nq, dcol - nacc)
#define FPNGB_FLUSHQ() do { \
        if (nq) { \
            uint32_t* b__ = reinterpret_cast<uint32_t*>(rowp + dcol - nacc - 4u * nq); \
            if (nq == 3u) { b__[0] = q1; b__[1] = q2; b__[2] = q3; } \
            else if (nq == 2u) { b__[0] = q2; b__[1] = q3; } \
            else b__[0] = q3; \
            nq = 0u; \
        } \
    } while (0)

    // general byte sink: unaligned head bytes of the thread's range and a scanline's last bytes go out singly, everything else
    // as aligned 32-bit words
#define FPNGB_EMIT(v) do { \
        const uint32_t v__ = (v); \
        if (nacc == 0u && (dcol & 3u)) rowp[dcol] = (uint8_t)v__; \
        else { acc |= (unsigned long long)v__ << (8u * nacc); if (++nacc == 4u) { *reinterpret_cast<uint32_t*>(rowp + dcol - 3u) = (uint32_t)acc; acc = 0ull; nacc = 0u; } } \
        if (++dcol == bpl) { \
            for (uint32_t i__ = 0; i__ < nacc; i__++) rowp[dcol - nacc + i__] = (uint8_t)(acc >> (8u * i__)); \
            acc = 0ull; nacc = 0u; row++; rowp += pitch; dcol = 0u; need_filter = true; \
        } \
    } while (0)
#define FPNGB_LITERAL(v) do { \
        const uint32_t s__ = (v); \
        lits = (lits >> 8) | (s__ << 24); \
        if (need_filter) { if (row >= h || s__ != (row ? 2u : 0u)) { *err = 1; return; } need_filter = false; } \
        else { if (row >= h) { *err = 1; return; } FPNGB_EMIT(s__); } \
    } while (0)

    // literal fast path allowed: inside a scanline (filter byte consumed, row valid) with the pending bytes word aligned; the fast path
    // itself keeps this true, every other path re-evaluates it when it is done
    bool fastok;
#define FPNGB_REFLAG() (fastok = !need_filter && row < h && (nacc | ((dcol & 3u) == 0u)))
    FPNGB_REFLAG();
    FastTok t;
    const uint32_t sizes_s = (uint32_t)__cvta_generic_to_shared(s_fast + 4096);
    while (rel < hi) {
        uint32_t run;
        if (fast_tok(win, s_fast, sizes_s, O + rel, rel, hi, t)) {
            rel += t.L;
            if (t.cnt) {
                const uint32_t cnt = t.cnt, P = t.payload;
                // fast path (almost every literal): inside a scanline, the pending bytes word aligned -- append the one to three bytes
                // to the 64-bit accumulator, store a word when four are there; no per-byte branches
                if (fastok && dcol + cnt < bpl) {
                    lits = __funnelshift_r(lits, P, 8u * cnt);
                    acc |= (unsigned long long)P << (8u * nacc);
                    nacc += cnt; dcol += cnt;
                    if (nacc >= 4u) {
                        const uint32_t wv = (uint32_t)acc;
                        acc >>= 32; nacc -= 4u;
                        const uint32_t wcol = dcol - nacc - 4u;              // data column of this word
                        if (nq == 0u && (wcol & 15u)) *reinterpret_cast<uint32_t*>(rowp + wcol) = wv;   // not yet at a 16-byte boundary
                        else {
                            q0 = q1; q1 = q2; q2 = q3; q3 = wv;
                            if (++nq == 4u) { *reinterpret_cast<uint4*>(rowp + wcol - 12u) = make_uint4(q0, q1, q2, q3); nq = 0u; }
                        }
                    }
                    continue;
                }
                FPNGB_FLUSHQ();
                FPNGB_LITERAL(P & 0xFFu);
                if (cnt > 1u) FPNGB_LITERAL((P >> 8) & 0xFFu);
                if (cnt > 2u) FPNGB_LITERAL(P >> 16);
                FPNGB_REFLAG();
                continue;
            }
            run = t.payload;
        } else {
            uint32_t sym;
            const uint32_t l = slow_tok(lutg, t.w, sym, run);
            if (!l) return;                                                // invalid code: the link pass already flagged it
            if (sym == 256u) break;
            rel += l;
            if (sym < 256u) { FPNGB_FLUSHQ(); FPNGB_LITERAL(sym); FPNGB_REFLAG(); continue; }
        }
        {
            FPNGB_FLUSHQ();
            const bool bad = row >= h || need_filter || dcol < (uint32_t)CHANS || (dcol % CHANS) != 0u || (run % CHANS) != 0u || dcol + run > bpl;
            if (bad) { *err = 1; return; }
            const uint32_t px = CHANS == 4 ? lits : (lits >> 8);          // last CHANS literals, oldest in the low byte
            if (CHANS == 4) {
                // dcol % 4 == 0: nothing is pending, pixels are word aligned
                uint32_t* d = reinterpret_cast<uint32_t*>(rowp + dcol);
                const uint32_t npx = run >> 2;
                uint32_t i = 0;
                for (; i < npx && ((dcol + 4u * i) & 15u); i++) d[i] = px;             // up to the next 16-byte boundary
                const uint4 px4 = make_uint4(px, px, px, px);
                for (; i + 4u <= npx; i += 4u) *reinterpret_cast<uint4*>(d + i) = px4;
                for (; i < npx; i++) d[i] = px;
                dcol += run;
                if (dcol == bpl) { row++; rowp += pitch; dcol = 0u; need_filter = true; }
            } else {
                // RGB: the 3-byte pixel repeats with a period of three 32-bit words.  Bytes up to the next word boundary go through
                // the byte sink, whole words are stored from three rotating pattern registers, the remainder again byte-wise.
                uint32_t left = run, ph = 0;                                  // ph: index inside the pixel of the next byte
                while (left && ((dcol & 3u) || nacc)) { FPNGB_EMIT((px >> (8u * ph)) & 0xFFu); ph = ph == 2u ? 0u : ph + 1u; left--; }
                if (left >= 4u) {
                    uint32_t wa = 0, wb = 0, wc = 0;
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++) {
                        wa |= ((px >> (8u * ((ph + j) % 3u))) & 0xFFu) << (8u * j);
                        wb |= ((px >> (8u * ((ph + 4u + j) % 3u))) & 0xFFu) << (8u * j);
                        wc |= ((px >> (8u * ((ph + 8u + j) % 3u))) & 0xFFu) << (8u * j);
                    }
                    const uint32_t nwords = left >> 2;
                    uint32_t* d = reinterpret_cast<uint32_t*>(rowp + dcol);
                    uint32_t k = 0;
                    for (; k < nwords && ((dcol + 4u * k) & 15u); k++) { d[k] = wa; const uint32_t tw = wa; wa = wb; wb = wc; wc = tw; }
                    for (; k + 4u <= nwords; k += 4u) {                       // four words advance the 3-word pattern by one
                        *reinterpret_cast<uint4*>(d + k) = make_uint4(wa, wb, wc, wa);
                        const uint32_t tw = wa; wa = wb; wb = wc; wc = tw;
                    }
                    for (; k < nwords; k++) { d[k] = wa; const uint32_t tw = wa; wa = wb; wb = wc; wc = tw; }
                    dcol += nwords << 2; left -= nwords << 2; ph = (ph + nwords) % 3u;      // 4 bytes advance the phase by 1
                    if (dcol == bpl) { row++; rowp += pitch; dcol = 0u; need_filter = true; }   // a run may end exactly at the end of its scanline
                }
                while (left) { FPNGB_EMIT((px >> (8u * ph)) & 0xFFu); ph = ph == 2u ? 0u : ph + 1u; left--; }
            }
            FPNGB_REFLAG();
        }
    }
    FPNGB_FLUSHQ();
    for (uint32_t i = 0; i < nacc; i++) rowp[dcol - nacc + i] = (uint8_t)(acc >> (8u * i));
#undef FPNGB_FLUSHQ
#undef FPNGB_REFLAG
#undef FPNGB_EMIT
#undef FPNGB_LITERAL
}

// "keep the last 4 literal bytes" monoid: a then b  (v holds the last n literals, most recent in the top byte)
__device__ __forceinline__ void lit_combine(uint32_t& n, uint32_t& v, uint32_t nb, uint32_t vb)
{
    if (nb >= 4u) { n = 4u; v = vb; return; }
    if (nb == 0u) return;
    v = (v >> (8u * nb)) | (vb & (0xFFFFFFFFu << (8u * (4u - nb))));
    n = min(n + nb, 4u);
}

// subsequence g of a file covers aligned-stream bits [g*kSubBits, (g+1)*kSubBits)
struct FileSpan { unsigned long long tok0; unsigned long long g0, g1; };     // first token bit; first/last subsequence index
__device__ __forceinline__ FileSpan file_span(const DecodeState& st, const Stream& sm, const FileDesc& fd)
{
    FileSpan f;
    f.tok0 = st.token_start + sm.bit0;
    const unsigned long long endbit = (unsigned long long)(fd.idat_len > 4u ? fd.idat_len - 4u : 0u) * 8ull + sm.bit0;   // tokens (and EOB) end before the Adler bytes
    f.g0 = f.tok0 / kSubBits;
    f.g1 = endbit ? (endbit - 1ull) / kSubBits : 0ull;
    if (f.g1 < f.g0) f.g1 = f.g0;
    return f;
}

__global__ void __launch_bounds__(kDecThreads) decode_scan_kernel(DecodeParams p)
{
    extern __shared__ __align__(16) uint32_t dec_smem[];
    uint32_t* s_fast = dec_smem;                      // [kFastWords]
    uint32_t* s_win = dec_smem + kFastWords;          // [kWinSmemWords]
    const uint32_t f = blockIdx.y, tid = threadIdx.x;
    const DecodeState st = p.state[f];
    if (st.status || st.stored) return;
    const FileDesc fd = p.files[f];
    Stream sm = open_stream(p.d_files + (size_t)f * p.file_stride, fd);
    const FileSpan sp = file_span(st, sm, fd);
    const unsigned long long g = sp.g0 + (unsigned long long)blockIdx.x * kDecThreads + tid;
    if (sp.g0 + (unsigned long long)blockIdx.x * kDecThreads > sp.g1) return;
    stage_fast_table(s_fast, p.fast + (size_t)f * kFastWords);
    stage_window(sm, s_win, sp.g0 + (unsigned long long)blockIdx.x * kDecThreads);
    stage_wait();
    if (g > sp.g1) return;
    const uint32_t* lutg = p.luts + (size_t)f * 4096;
    SubInfo* info = p.subs + (size_t)f * p.subs_per_file + (g - sp.g0);
    const unsigned long long lo_abs = g * kSubBits;
    SubScan r;
    unsigned long long origin;
    if (g == sp.g0) {            // the file's first subsequence starts exactly at the first token
        origin = sp.tok0;
        r = scan_window(s_win, (uint32_t)(origin - 32ull * sm.sw_base), s_fast, lutg, 0u, 0u, (uint32_t)((g + 1) * kSubBits - origin));
    } else {
        origin = lo_abs - kPreRoll < sp.tok0 ? sp.tok0 : lo_abs - kPreRoll;     // never pre-roll across the block header
        r = scan_window(s_win, (uint32_t)(origin - 32ull * sm.sw_base), s_fast, lutg, 0u, (uint32_t)(lo_abs - origin), (uint32_t)(lo_abs - origin) + kSubBits);
    }
    info->start = origin + r.first;
    info->exit = r.exit >= kRelErr ? (r.exit == kRelEnd ? kPosEnd : kPosErr) : origin + r.exit;
    info->eob_end = origin + r.eob_end;
    info->n_out = r.n_out; info->lits = r.lits; info->nlit = min(r.nlit, 4u);
}

// statistics: subsequences whose speculative start was wrong and had to be decoded again by the link pass (fpngb_debug_decode_repairs)
__device__ unsigned long long g_link_repairs = 0ull;
unsigned long long decode_link_repairs(bool reset)
{
    unsigned long long v = 0, z = 0;
    cudaMemcpyFromSymbol(&v, g_link_repairs, sizeof v);
    if (reset) cudaMemcpyToSymbol(g_link_repairs, &z, sizeof z);
    return v;
}

__global__ void __launch_bounds__(kLinkThreads) decode_link_kernel(DecodeParams p)
{
    __shared__ uint32_t s_lut[4096];
    __shared__ unsigned long long s_exit[kLinkThreads];
    __shared__ unsigned long long s_scan[kLinkThreads / 32];
    __shared__ uint32_t s_litn[kLinkThreads / 32], s_litv[kLinkThreads / 32];
    __shared__ unsigned long long s_wbase[kLinkThreads / 32], s_chunk_total;
    __shared__ uint32_t s_wlitn[kLinkThreads / 32], s_wlitv[kLinkThreads / 32], s_chunk_litn, s_chunk_litv;
    __shared__ unsigned long long s_carry_start, s_carry_out, s_eob_end;
    __shared__ uint32_t s_carry_litn, s_carry_litv, s_done, s_err, s_first_bad, s_lut_loaded;

    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DecodeState* stp = p.state + f;
    if (stp->status || stp->stored) return;
    const DecodeState st = *stp;
    const FileDesc fd = p.files[f];
    const Stream sm = open_stream(p.d_files + (size_t)f * p.file_stride, fd);
    const FileSpan sp = file_span(st, sm, fd);
    const uint32_t chans = p.chans, bpl = p.w * chans, h = p.h;
    const unsigned long long total_out = (unsigned long long)(bpl + 1) * h;
    SubInfo* subs = p.subs + (size_t)f * p.subs_per_file;
    const unsigned long long nsub = sp.g1 - sp.g0 + 1ull;

    if (tid == 0) { s_carry_start = sp.tok0; s_carry_out = 0; s_carry_litn = 0; s_carry_litv = 0; s_done = 0; s_err = 0; s_eob_end = 0; s_lut_loaded = 0; }
    __syncthreads();

    SubInfo nxt;                                       // the next chunk's records are loaded while this chunk is processed
    if (tid < nsub) nxt = subs[tid];
    for (unsigned long long base = 0; base < nsub; base += kLinkThreads) {
        const unsigned long long i = base + tid;
        const bool have = i < nsub;
        SubInfo in;
        if (have) in = nxt;
        else { in.start = 0; in.exit = kPosErr; in.eob_end = 0; in.n_out = 0; in.lits = 0; in.nlit = 0; }
        if (i + kLinkThreads < nsub) nxt = subs[i + kLinkThreads];
        const unsigned long long g = sp.g0 + i;
        // ---- verify / repair the chain: every subsequence must start where its predecessor exits
        for (uint32_t iter = 0;; iter++) {
            s_exit[tid] = in.exit;
            __syncthreads();
            const unsigned long long want = tid == 0 ? s_carry_start : s_exit[tid - 1];
            const int mismatch = have && want < kPosErr && want != in.start;
            if (!__syncthreads_or(mismatch)) break;
            if (!s_lut_loaded) {                       // repairs are rare: fetch the LUT only when first needed
                for (uint32_t k = tid; k < 4096; k += blockDim.x) s_lut[k] = p.luts[(size_t)f * 4096 + k];
                __syncthreads();
                if (tid == 0) s_lut_loaded = 1;
            }
            if (mismatch) {
                atomicAdd(&g_link_repairs, 1ull);
                const unsigned long long hi_abs = (g + 1) * kSubBits;
                if (want >= hi_abs) { in.start = want; in.exit = want; in.n_out = 0; in.nlit = 0; in.lits = 0; }   // cannot happen for kSubBits >> token size
                else {
                    const SubScan r = decode_range<false>(sm, s_lut, want, 0u, 0u, (uint32_t)(hi_abs - want), chans, nullptr, 0, 0, 0, 0, 0, nullptr);
                    in.start = want;
                    in.exit = r.exit >= kRelErr ? (r.exit == kRelEnd ? kPosEnd : kPosErr) : want + r.exit;
                    in.eob_end = want + r.eob_end; in.n_out = r.n_out; in.lits = r.lits; in.nlit = min(r.nlit, 4u);
                }
            }
            __syncthreads();
        }
        // the first subsequence that ends in EOB / an invalid code terminates the stream; everything after it is dead
        if (tid == 0) s_first_bad = kLinkThreads;
        __syncthreads();
        if (in.exit >= kPosErr) atomicMin(&s_first_bad, tid);
        __syncthreads();
        const uint32_t first_bad = s_first_bad;
        const bool live = have && tid <= first_bad;
        if (!live) { in.n_out = 0; in.nlit = 0; in.lits = 0; }

        // ---- block-exclusive scans: output byte offset and literal window entering each subsequence
        unsigned long long incl = in.n_out;
        uint32_t wn = in.nlit, wv = in.lits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(kFullMask, incl, o);
            const uint32_t un = __shfl_up_sync(kFullMask, wn, o), uv = __shfl_up_sync(kFullMask, wv, o);
            if (lane >= (uint32_t)o) { incl += up; uint32_t tn = un, tv = uv; lit_combine(tn, tv, wn, wv); wn = tn; wv = tv; }
        }
        if (lane == 31) { s_scan[warp] = incl; s_litn[warp] = wn; s_litv[warp] = wv; }
        unsigned long long ex = __shfl_up_sync(kFullMask, incl, 1);
        uint32_t en = __shfl_up_sync(kFullMask, wn, 1), ev = __shfl_up_sync(kFullMask, wv, 1);
        if (lane == 0) { ex = 0; en = 0; ev = 0; }
        __syncthreads();
        // second level: warp 0 scans the (<= 32) warp totals, carry of the earlier chunks folded in
        static_assert(kLinkThreads / 32 <= 32, "one warp scans the warp totals");
        if (warp == 0) {
            const bool hw = lane < kLinkThreads / 32;
            unsigned long long t = hw ? s_scan[lane] : 0ull;
            uint32_t tn = hw ? s_litn[lane] : 0u, tv = hw ? s_litv[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(kFullMask, t, o);
                const uint32_t un = __shfl_up_sync(kFullMask, tn, o), uv = __shfl_up_sync(kFullMask, tv, o);
                if (lane >= (uint32_t)o) { t += up; uint32_t an = un, av = uv; lit_combine(an, av, tn, tv); tn = an; tv = av; }
            }
            // exclusive form, then the carry in front
            unsigned long long xt = __shfl_up_sync(kFullMask, t, 1);
            uint32_t xn = __shfl_up_sync(kFullMask, tn, 1), xv = __shfl_up_sync(kFullMask, tv, 1);
            if (lane == 0) { xt = 0; xn = 0; xv = 0; }
            uint32_t bn0 = s_carry_litn, bv0 = s_carry_litv;
            lit_combine(bn0, bv0, xn, xv);
            if (hw) { s_wbase[lane] = xt; s_wlitn[lane] = bn0; s_wlitv[lane] = bv0; }
            if (lane == 31) {                                             // lanes beyond the last warp carry zeros: lane 31 holds the chunk totals
                uint32_t cn0 = s_carry_litn, cv0 = s_carry_litv;
                lit_combine(cn0, cv0, tn, tv);
                s_chunk_total = t; s_chunk_litn = cn0; s_chunk_litv = cv0;
            }
        }
        __syncthreads();
        const unsigned long long wbase = s_wbase[warp], chunk_total = s_chunk_total;
        uint32_t bn = s_wlitn[warp], bv = s_wlitv[warp];
        const uint32_t cn = s_chunk_litn, cv = s_chunk_litv;
        lit_combine(bn, bv, en, ev);
        const unsigned long long out_pos = s_carry_out + wbase + ex;

        if (have) {
            SubInfo o;
            o.start = in.start; o.exit = out_pos; o.eob_end = 0; o.n_out = live ? in.n_out : 0u; o.lits = bv; o.nlit = live ? 1u : 0u;
            if (live && out_pos + in.n_out > total_out) { o.n_out = 0; o.nlit = 0; s_err = 1; }
            subs[i] = o;
        }
        if (have && tid == first_bad) { if (in.exit == kPosEnd) { s_eob_end = in.eob_end; s_done = 1; } else s_err = 1; }
        __syncthreads();
        if (tid == 0) {
            s_carry_start = s_exit[kLinkThreads - 1];
            s_carry_out += chunk_total; s_carry_litn = cn; s_carry_litv = cv;
            if (s_carry_out > total_out) s_err = 1;
        }
        __syncthreads();
        if (s_done || s_err) {
            // mark the remaining subsequences dead so the write kernel skips them
            for (unsigned long long k = base + kLinkThreads + tid; k < nsub; k += kLinkThreads) { subs[k].n_out = 0; subs[k].nlit = 0; }
            break;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // EOB right after the last row, then pad to a byte, then exactly the 4 Adler bytes (fpng.cpp:2559-2584)
        const unsigned long long used = s_eob_end >= sm.bit0 ? ((s_eob_end - sm.bit0 + 7ull) >> 3) : 0ull;
        const bool ok = !s_err && s_done && s_carry_out == total_out && used + 4ull == fd.idat_len;
        stp->status = ok ? 0u : 1u;
        stp->out_bytes = s_carry_out; stp->end_byte = used;
    }
}

__global__ void __launch_bounds__(kDecThreads) decode_write_kernel(DecodeParams p)
{
    extern __shared__ __align__(16) uint32_t dec_smem[];
    uint32_t* s_fast = dec_smem;                      // [kFastWords]
    uint32_t* s_win = dec_smem + kFastWords;          // [kWinSmemWords]
    const uint32_t f = blockIdx.y, tid = threadIdx.x;
    DecodeState* stp = p.state + f;
    if (stp->status || stp->stored) return;
    const FileDesc fd = p.files[f];
    Stream sm = open_stream(p.d_files + (size_t)f * p.file_stride, fd);
    const FileSpan sp = file_span(*stp, sm, fd);
    if (sp.g0 + (unsigned long long)blockIdx.x * kDecThreads > sp.g1) return;
    const unsigned long long g = sp.g0 + (unsigned long long)blockIdx.x * kDecThreads + tid;
    SubInfo in; in.nlit = 0; in.n_out = 0;
    if (g <= sp.g1) in = p.subs[(size_t)f * p.subs_per_file + (g - sp.g0)];
    // decode_write_staged_kernel ran first and cleared the live flag of everything it wrote (all of it, for literal-dominated files)
    if (!__syncthreads_or(in.nlit && in.n_out)) return;
    stage_fast_table(s_fast, p.fast + (size_t)f * kFastWords);
    stage_window(sm, s_win, sp.g0 + (unsigned long long)blockIdx.x * kDecThreads);
    stage_wait();
    const uint32_t* lutg = p.luts + (size_t)f * 4096;
    if (g > sp.g1) return;
    if (!in.nlit || !in.n_out) return;                                           // dead, empty or already written
    const uint32_t chans = p.chans, bpl = p.w * chans, h = p.h, pitch = p.delta_pitch;
    uint32_t err = 0;
    const unsigned long long hi_abs = (g + 1) * kSubBits;
    if (in.start < hi_abs) {
        uint8_t* dl = p.delta + (size_t)f * pitch * h;
        if (chans == 4) decode_write_range<4>(sm, s_fast, lutg, in.start, (uint32_t)(hi_abs - in.start), dl, pitch, bpl, h, in.exit /*out_pos*/, in.lits /*tail*/, &err);
        else decode_write_range<3>(sm, s_fast, lutg, in.start, (uint32_t)(hi_abs - in.start), dl, pitch, bpl, h, in.exit /*out_pos*/, in.lits /*tail*/, &err);
    }
    if (err) stp->status = 1;
}

// ------------------------------------------------------------------------------------------------
// D2c' decode_write_staged_kernel: the write pass for literal-dominated streams (photographic content; runs first, the kernel above
// takes what it leaves).  ncu of the kernel above: ~100 warp instructions per loop iteration for ~2.2 bytes per lane -- every lane
// is somewhere else in its scanline, so the warp executes the union of the byte sink's paths (word complete / queue / 128-bit
// store / scanline end) in almost every iteration, and every lane's stores go to a cache line of their own.  Here the CTA's output
// (a contiguous range of the filtered stream, filter bytes included) is assembled in shared memory first:
//   * the loop body is straight-line: table look-up, append the 1-3 literal bytes to a 64-bit accumulator, and when a 32-bit word is
//     complete OR it into the staging buffer (predicated red.shared.or; the words two neighbouring subsequences share need no
//     special case because the buffer starts zeroed).  No scanline bookkeeping: a match computes its column from its stream position
//     when it needs it (fpng.cpp:2289-2388 checks), the filter bytes are checked by the copy below (fpng.cpp:2253-2262);
//   * afterwards the CTA copies the staged bytes to the delta rows (filter bytes dropped, scanlines at their pitch) with realigning
//     128-bit stores: coalesced, one transaction per 16 bytes instead of one per lane and word.
// A CTA whose subsequences produce more than kWsStageBytes (RLE-dominated content) leaves them to decode_write_kernel; the ones
// written here get their live flag (SubInfo::nlit) cleared.  Statuses and pixels are identical either way (same tokens, same checks).
// ------------------------------------------------------------------------------------------------
constexpr int kWsThreads = 224;                        // subsequences per CTA (2 CTAs per SM with the 64 KiB staging buffer)
constexpr uint32_t kWsStageBytes = 64u * 1024u;
constexpr uint32_t kWsWinWords = kWinLead + kWsThreads * (kSubBits / 32) + kWinTail;
constexpr uint32_t kWsWinSmemWords = (kWsWinWords + kWsWinWords / 32 + 1 + 3) / 4 * 4;
constexpr uint32_t kWsStageWords = kWsStageBytes / 4 + 8;                         // slack: the realigning copy reads one word past the data
constexpr size_t kWsSmem = (size_t)(kFastWords + kWsWinSmemWords + kWsStageWords) * 4;

__device__ __forceinline__ void stage_window_n(Stream& st, uint32_t* s_win, unsigned long long first_sub, uint32_t nwords)
{
    const unsigned long long w0 = first_sub * (kSubBits / 32);
    const uint32_t base = (uint32_t)(w0 >= kWinLead ? w0 - kWinLead : 0ull);
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) {                 // same layout as stage_window()
        const uint32_t* src = st.words + min(base + i, st.max_widx);
        cp_async4(s_win + i + (i >> 5), src);
        if ((i & 31u) == 0u && i) cp_async4(s_win + i + (i >> 5) - 1u, src);
    }
    st.swin = s_win; st.sw_base = base; st.sw_count = nwords;
}

// A value the compiler must keep in its register: shared-window addresses are otherwise re-derived (S2R SR_CgaCtaId + LEA) inside the
// decode loop, on its critical path.
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { uint32_t r; asm volatile("mov.b32 %0, %1;\n" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(saddr)); return v; }
// win_peek() on the window's shared-window address
__device__ __forceinline__ uint32_t win_peek_s(uint32_t win_s, uint32_t q)
{
    const uint32_t si = q >> 5, a = win_s + ((si + (si >> 5)) << 2);
    return __funnelshift_r(lds32(a), lds32(a + 4u), q);
}

// Byte sink of the staged write pass: `lo`/`hi` hold the pending bytes (sh / 8 < 4 of them, + up to 4 new ones) of the word at shared
// address `saddr`; sh = 8 * pending bytes.
struct StageSink {
    uint32_t lo, hi, sh, saddr, send;
    __device__ __forceinline__ void begin(uint32_t stage_s, uint32_t byte_ofs, uint32_t stage_end_s)
    {
        lo = 0u; hi = 0u; sh = 8u * (byte_ofs & 3u); saddr = stage_s + (byte_ofs & ~3u); send = stage_end_s;
    }
    // append the bits8 / 8 (1..4) low bytes of v (higher bytes zero)
    __device__ __forceinline__ void put(uint32_t v, uint32_t bits8)
    {
        lo |= v << sh;
        hi |= __funnelshift_l(v, 0u, sh);                                  // the bytes that spill into the next word (0 when sh == 0)
        sh += bits8;
        // word complete: plain store (this thread wrote the word's last byte, so it is the only one that ever stores it; bytes of the
        // word that belong to a neighbouring subsequence are zero here and are OR-ed in after the CTA barrier, see tail());
        // never beyond the buffer, whatever the stream holds
        asm volatile("{\n.reg .pred p;\nsetp.ge.u32 p, %1, 32;\nsetp.lt.and.u32 p, %0, %3, p;\n@p st.shared.u32 [%0], %2;\n}\n"
                     :: "r"(saddr), "r"(sh), "r"(lo), "r"(send) : "memory");
        const bool adv = sh >= 32u;
        saddr += adv ? 4u : 0u;
        lo = adv ? hi : lo;
        hi = adv ? 0u : hi;
        sh &= 31u;
    }
    // the last, incomplete word: OR-ed into the buffer AFTER every thread of the CTA has finished its stores (shared atomics in the
    // loop itself made the LSU the bottleneck: measured 3.0 ms vs 2.4 ms for the per-thread sink on C2)
    __device__ __forceinline__ void tail() const
    {
        if (lo && saddr < send) asm volatile("red.shared.or.b32 [%0], %1;\n" :: "r"(saddr), "r"(lo) : "memory");
    }
    __device__ __forceinline__ uint32_t pos(uint32_t stage_s) const { return saddr - stage_s + (sh >> 3); }   // byte offset of the next byte
};

// General step of the staged write loop: one token the way decode_write_range takes it (trimming at the end of the subsequence, matches,
// slow-table tokens, end of block).  Returns true when the subsequence is finished.
template <int CHANS>
__device__ __forceinline__ bool stage_general_step(StageSink& sk, const uint32_t* __restrict__ win, uint32_t O, const uint32_t* __restrict__ s_fast, uint32_t sizes_s,
                                                const uint32_t* __restrict__ lutg, uint32_t hi, uint32_t stage_s, uint32_t byte_ofs, uint32_t col0, uint32_t bpl,
                                                uint32_t& rel, uint32_t& lits, uint32_t* err)
{
    if (rel >= hi) return true;
    FastTok t;
    uint32_t run;
    if (fast_tok(win, s_fast, sizes_s, O + rel, rel, hi, t)) {
        rel += t.L;
        if (t.cnt) {
            lits = __funnelshift_r(lits, t.payload, 8u * t.cnt);
            sk.put(t.payload, 8u * t.cnt);
            return false;
        }
        run = t.payload;
    } else {
        uint32_t sym;
        const uint32_t l = slow_tok(lutg, t.w, sym, run);
        if (!l) return true;                                                 // invalid code: the link pass already flagged it
        if (sym == 256u) return true;
        rel += l;
        if (sym < 256u) { lits = (lits >> 8) | (sym << 24); sk.put(sym, 8u); return false; }
    }
    // RLE match (fpng.cpp:2289-2388): inside a scanline, pixel aligned, not across its end; replicates the previous delta pixel
    const uint32_t c = (col0 + (sk.pos(stage_s) - byte_ofs)) % (bpl + 1u);
    const uint32_t dcol = c - 1u;
    if (c == 0u || dcol < (uint32_t)CHANS || (dcol % CHANS) != 0u || (run % CHANS) != 0u || dcol + run > bpl) { *err = 1; return true; }
    const uint32_t px = CHANS == 4 ? lits : (lits >> 8);                      // last CHANS literals, oldest in the low byte
    for (uint32_t i = 0; i < run; i += CHANS) sk.put(px, 8u * CHANS);
    return false;
}

// One subsequence into the staging buffer.  `col0`: column of its first output byte in the filtered stream (0 = filter byte).
// The loop is software-pipelined around its only true dependency (token length -> next bit position -> next window read -> next table
// entry): the window read of the NEXT token is issued as soon as this token's length is known, before its bytes go to the sink; the
// common case (1-3 literals, more than 12 bits away from the end of the subsequence, so no trimming can apply) is one straight-line
// block behind one branch.  Everything else takes the general step, which is the loop of decode_write_range.
template <int CHANS>
__device__ __forceinline__ void decode_stage_range(uint32_t mask, StageSink& sk, const uint32_t* __restrict__ win, uint32_t O, const uint32_t* __restrict__ s_fast, const uint32_t* __restrict__ lutg,
                                                   uint32_t hi, uint32_t stage_s, uint32_t stage_end_s, uint32_t byte_ofs, uint32_t col0, uint32_t bpl,
                                                   uint32_t tail, uint32_t* err)
{
    sk.begin(stage_s, byte_ofs, stage_end_s);
    uint32_t rel = 0, lits = tail;
    const uint32_t sizes_s = (uint32_t)__cvta_generic_to_shared(s_fast + 4096);
    const uint32_t fast_s = pin_u32((uint32_t)__cvta_generic_to_shared(s_fast)), win_s = pin_u32((uint32_t)__cvta_generic_to_shared(win));
    uint32_t w = win_peek_s(win_s, O);
    // The warp votes once per iteration, which makes every iteration a convergence point: left to itself the compiler peels the
    // literal block into an inner loop, and a lane that needs the general step (a match, ~1 in 450 tokens of photographic content) then
    // waits until EVERY other lane of the warp needs one too -- ncu of the first version: 20 of 32 lanes active in the literal block.
    // The end of the subsequence (a multi-literal entry cut down to its first literal within 12 bits of the end, same rule as
    // fast_tok) is handled inside the literal block with predicated instructions, so the general step stays rare.
    while (true) {
        const bool more = rel < hi;
        if (!__any_sync(mask, more)) break;
        if (more) {
            const uint32_t e = lds32(fast_s + ((w & 4095u) << 2));
            uint32_t L = e & 15u, c8 = (e >> 1) & 0x18u, P = e >> 8;        // c8 = 8 * literal count
            if (L != 0u && c8 != 0u) {
                if (c8 >= 16u && rel + 12u > hi) {
                    P &= 0xFFu; c8 = 8u;
                    asm volatile("ld.shared.u8 %0, [%1+16384];\n" : "=r"(L) : "r"(fast_s + P));   // code size of the first literal
                }
                rel += L;
                w = win_peek_s(win_s, O + rel);
                lits = __funnelshift_r(lits, P, c8);
                sk.put(P, c8);
            } else {
                if (stage_general_step<CHANS>(sk, win, O, s_fast, sizes_s, lutg, hi, stage_s, byte_ofs, col0, bpl, rel, lits, err)) rel = hi;
                w = win_peek_s(win_s, O + rel);
            }
        }
    }
}

__global__ void __launch_bounds__(kWsThreads) decode_write_staged_kernel(DecodeParams p)
{
    extern __shared__ __align__(16) uint32_t dec_smem[];
    uint32_t* s_fast = dec_smem;                            // [kFastWords]
    uint32_t* s_win = dec_smem + kFastWords;                // [kWsWinSmemWords]
    uint32_t* s_stage = s_win + kWsWinSmemWords;            // [kWsStageWords]
    __shared__ unsigned long long s_lo, s_hi, s_wlo[kWsThreads / 32], s_whi[kWsThreads / 32];
    __shared__ uint32_t s_err;
    const uint32_t f = blockIdx.y, tid = threadIdx.x;
    DecodeState* stp = p.state + f;
    if (stp->status || stp->stored) return;
    const FileDesc fd = p.files[f];
    Stream sm = open_stream(p.d_files + (size_t)f * p.file_stride, fd);
    const FileSpan sp = file_span(*stp, sm, fd);
    const unsigned long long first_sub = sp.g0 + (unsigned long long)blockIdx.x * kWsThreads;
    if (first_sub > sp.g1) return;
    const unsigned long long g = first_sub + tid;
    SubInfo* sub = p.subs + (size_t)f * p.subs_per_file + (g - sp.g0);
    SubInfo in; in.nlit = 0; in.n_out = 0; in.start = 0; in.exit = 0; in.lits = 0;
    if (g <= sp.g1) in = *sub;
    const unsigned long long hi_abs = (g + 1) * kSubBits;
    const bool live = in.nlit && in.n_out && in.start < hi_abs;
    // the CTA's output range in the filtered stream: warp reductions, then one thread combines the warps
    {
        unsigned long long vlo = live ? in.exit : ~0ull, vhi = live ? in.exit + in.n_out : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(kFullMask, vlo, o), b = __shfl_xor_sync(kFullMask, vhi, o);
            vlo = a < vlo ? a : vlo; vhi = b > vhi ? b : vhi;
        }
        if ((tid & 31u) == 0u) { s_wlo[tid >> 5] = vlo; s_whi[tid >> 5] = vhi; }
        if (tid == 0) s_err = 0u;
        __syncthreads();
        if (tid == 0) {
            for (uint32_t k = 1; k < kWsThreads / 32; k++) { vlo = s_wlo[k] < vlo ? s_wlo[k] : vlo; vhi = s_whi[k] > vhi ? s_whi[k] : vhi; }
            s_lo = vlo; s_hi = vhi;
        }
        __syncthreads();
    }
    const unsigned long long lo = s_lo, hi = s_hi;
    if (lo >= hi) return;                                                        // nothing to write
    const unsigned long long bw = lo & ~15ull;                                   // stream offset of staging byte 0
    if (hi - bw > (unsigned long long)kWsStageBytes) return;                     // too much output to stage: decode_write_kernel
    stage_fast_table(s_fast, p.fast + (size_t)f * kFastWords);
    stage_window_n(sm, s_win, first_sub, kWsWinWords);
    // only the range's last (possibly incomplete) word and the slack the realigning copy reads beyond it need zeroing: every other
    // word is stored whole by the thread that completes it
    if (tid < 8u) s_stage[(uint32_t)((hi - bw) >> 2) + tid] = 0u;
    stage_wait();

    const uint32_t chans = p.chans, bpl = p.w * chans, h = p.h, pitch = p.delta_pitch, rb = bpl + 1u;
    StageSink sk; sk.lo = 0u; sk.saddr = 0u; sk.send = 0u;
    const uint32_t live_mask = __ballot_sync(kFullMask, live);                   // the lanes that decode (they vote together inside)
    if (live) {
        const uint32_t* lutg = p.luts + (size_t)f * 4096;
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(s_stage);
        const uint32_t O = (uint32_t)(in.start - 32ull * sm.sw_base);
        const uint32_t col0 = (uint32_t)(in.exit % rb);
        uint32_t err = 0;
        if (chans == 4) decode_stage_range<4>(live_mask, sk, s_win, O, s_fast, lutg, (uint32_t)(hi_abs - in.start), stage_s, stage_s + kWsStageBytes, (uint32_t)(in.exit - bw), col0, bpl, in.lits, &err);
        else decode_stage_range<3>(live_mask, sk, s_win, O, s_fast, lutg, (uint32_t)(hi_abs - in.start), stage_s, stage_s + kWsStageBytes, (uint32_t)(in.exit - bw), col0, bpl, in.lits, &err);
        if (err) s_err = 1u;
        sub->nlit = 0u;                                                          // written (or failed): nothing left for decode_write_kernel
    }
    __syncthreads();
    sk.tail();
    __syncthreads();

    // ---- staged stream bytes [lo, hi) -> delta rows
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_stage);
    uint8_t* dl = p.delta + (size_t)f * pitch * h;
    const unsigned long long y0 = lo / rb, y1 = (hi - 1ull) / rb;                // scanlines the range touches (link pass: hi <= rb * h)
    uint32_t bad = 0;
    for (unsigned long long y = y0 + tid; y <= y1; y += kWsThreads) {            // filter bytes: 0 on the first scanline, 2 (Up) below (fpng.cpp:2253-2262)
        const unsigned long long s = y * rb;
        if (s >= lo && s < hi && sb[(uint32_t)(s - bw)] != (y ? 2u : 0u)) bad = 1u;
    }
    if (rb < 256u) {
        // narrow images: one thread per stream byte
        for (unsigned long long s = lo + tid; s < hi; s += kWsThreads) {
            const unsigned long long y = s / rb; const uint32_t c = (uint32_t)(s - y * rb);
            if (c) dl[(size_t)y * pitch + (c - 1u)] = sb[(uint32_t)(s - bw)];
        }
    } else {
        for (unsigned long long y = y0; y <= y1; y++) {                          // warp-uniform: a scanline per round
            const unsigned long long rs = y * rb + 1ull;                         // stream offset of the scanline's first data byte
            const unsigned long long a = rs > lo ? rs : lo, b = rs + bpl < hi ? rs + bpl : hi;
            if (a >= b) continue;
            const uint32_t d_lo = (uint32_t)(a - rs), d_hi = (uint32_t)(b - rs); // data columns [d_lo, d_hi) of this scanline are staged here
            uint8_t* drow = dl + (size_t)y * pitch;
            const uint32_t src0 = (uint32_t)(rs - bw);                           // staging byte offset of column 0 (may lie before the buffer: only
                                                                                 // columns >= d_lo are read); unsigned wrap-around is fine below
            for (uint32_t c = (d_lo >> 4) + tid; 16u * c < d_hi; c += kWsThreads) {
                const uint32_t c0 = 16u * c, c1 = min(c0 + 16u, bpl);
                if (c0 >= d_lo && c1 <= d_hi && c1 - c0 == 16u) {
                    const uint32_t o = src0 + c0, k = o >> 2, sh = 8u * (o & 3u);
                    const uint32_t w0 = s_stage[k], w1 = s_stage[k + 1], w2 = s_stage[k + 2], w3 = s_stage[k + 3], w4 = s_stage[k + 4];
                    *reinterpret_cast<uint4*>(drow + c0) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                                                      __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
                } else {
                    for (uint32_t x = max(c0, d_lo); x < min(c1, d_hi); x++) drow[x] = sb[src0 + x];
                }
            }
        }
    }
    if (bad) s_err = 1u;
    __syncthreads();
    if (tid == 0 && s_err) stp->status = 1;
}

// ------------------------------------------------------------------------------------------------
// D2': stored blocks -> delta rows (filter bytes must all be 0; "delta" then already holds final pixel bytes)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_stored_kernel(DecodeParams p)
{
    const uint32_t f = blockIdx.y;
    DecodeState* st = p.state + f;
    if (st->status || !st->stored) return;
    const FileDesc fd = p.files[f];
    const uint8_t* z = p.d_files + (size_t)f * p.file_stride + fd.idat_ofs + 8;
    const uint32_t avail = fd.file_size - (fd.idat_ofs + 8);
    const uint32_t bpl = p.w * p.chans, h = p.h, pitch = p.delta_pitch;
    const unsigned long long raw = (unsigned long long)(bpl + 1) * h, nblk = (raw + 65534ull) / 65535ull;
    uint8_t* delta = p.delta + (size_t)f * pitch * h;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t y = blockIdx.x * 8 + warp;
    bool bad = false;
    // block headers: every warp of block row 0 validates a strided share (fpng.cpp:2124-2146, 2195-2204)
    if (blockIdx.x == 0) {
        if (2ull + raw + 5ull * nblk + 4ull != fd.idat_len || fd.idat_len > avail) bad = true;
        else for (unsigned long long j = threadIdx.x; j < nblk; j += blockDim.x) {
            const uint8_t* b = z + 2ull + j * 65540ull;
            const unsigned long long remaining = raw - j * 65535ull;
            const uint32_t want = remaining < 65535ull ? (uint32_t)remaining : 65535u;
            const uint32_t len = b[1] | (b[2] << 8), nlen = b[3] | (b[4] << 8);
            const uint32_t fin = (j + 1 == nblk) ? 1u : 0u;
            if ((b[0] & 1u) != fin || ((b[0] >> 1) & 3u) != 0 || len != want || len != (~nlen & 0xFFFFu)) bad = true;
        }
    }
    if (y < h && !(2ull + raw + 5ull * nblk + 4ull != fd.idat_len || fd.idat_len > avail)) {
        const unsigned long long s0 = (unsigned long long)y * (bpl + 1ull);
        for (uint32_t t = lane; t <= bpl; t += 32) {
            const unsigned long long s = s0 + t;
            const uint8_t v = z[2ull + 5ull * (s / 65535ull + 1ull) + s];
            if (t == 0) { if (v != 0) bad = true; }                               // fpng.cpp:2154-2158
            else delta[(size_t)y * pitch + (t - 1)] = v;
        }
    }
    if (bad) st->status = 1;
}

// ------------------------------------------------------------------------------------------------
// D3: inverse Up filter (running sum down the columns) + channel conversion.  One thread per group of 4 pixels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vadd4(uint32_t a, uint32_t b)
{
    const uint32_t s = (a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu);
    return s ^ ((a ^ b) & 0x80808080u);
}

template <int SRC, int DST>
__global__ void __launch_bounds__(128) unfilter_kernel(DecodeParams p)
{
    const uint32_t f = blockIdx.y;
    const DecodeState st = p.state[f];
    if (st.status) return;
    const uint32_t w = p.w, h = p.h, pitch = p.delta_pitch;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;                   // pixel group index
    const uint32_t x0 = g * 4;
    const uint32_t amask = __ballot_sync(kFullMask, x0 < w);                     // lanes of this warp that own a group (they stay converged)
    if (x0 >= w) return;
    const uint32_t lane = threadIdx.x & 31u;
    const bool next_full = lane < 31u && x0 + 8u <= w;                           // the next lane of this warp owns a whole group
    const uint32_t npx = min(4u, w - x0);
    const uint8_t* delta = p.delta + (size_t)f * pitch * h + (size_t)x0 * SRC;
    uint8_t* out = p.d_out + (size_t)f * p.out_stride + (size_t)x0 * DST;
    const size_t out_pitch = (size_t)w * DST;
    const bool full = npx == 4;
    const bool out_aligned = (((uintptr_t)out | out_pitch) & 3) == 0;
    uint32_t acc[SRC];
#pragma unroll
    for (int i = 0; i < SRC; i++) acc[i] = 0;
    const bool summing = !st.stored;
    constexpr int kRowsAhead = 8;                  // independent loads in flight per thread (8 rows x 12-16 bytes: enough bytes in flight at 256 files)
    for (uint32_t y0 = 0; y0 < h; y0 += kRowsAhead) {
        uint32_t dd[kRowsAhead][SRC];
#pragma unroll
        for (int j = 0; j < kRowsAhead; j++) {
            const uint32_t y = min(y0 + j, h - 1);
            const uint8_t* dr = delta + (size_t)y * pitch;
            if (full) {
#pragma unroll
                for (int i = 0; i < SRC; i++) dd[j][i] = __ldg(reinterpret_cast<const uint32_t*>(dr + 4 * i));   // pitch % 16 == 0 and x0*SRC % 4 == 0
            } else {
#pragma unroll
                for (int i = 0; i < SRC; i++) { dd[j][i] = 0; for (int b = 0; b < 4; b++) if ((uint32_t)(4 * i + b) < npx * SRC) dd[j][i] |= (uint32_t)dr[4 * i + b] << (8 * b); }
            }
        }
#pragma unroll
        for (int j = 0; j < kRowsAhead; j++) {
            const uint32_t y = y0 + j;
            if (y >= h) break;
#pragma unroll
            for (int i = 0; i < SRC; i++) acc[i] = summing ? vadd4(acc[i], dd[j][i]) : dd[j][i];
            // repack to DST channels
            uint32_t o[DST];
            if (SRC == DST) {
#pragma unroll
                for (int i = 0; i < DST; i++) o[i] = acc[i];
            } else if (SRC == 3) {   // 24 -> 32 bpp, alpha 0xFF (fpng.cpp:2331)
                o[0] = (acc[0] & 0x00FFFFFFu) | 0xFF000000u;
                o[1] = __byte_perm(acc[0], acc[1], 0x4543) | 0xFF000000u;
                o[2] = __byte_perm(acc[1], acc[2], 0x4432) | 0xFF000000u;
                o[DST - 1] = (acc[2] >> 8) | 0xFF000000u;
            } else {                 // 32 -> 24 bpp, alpha dropped (fpng.cpp:2684-2721)
                o[0] = __byte_perm(acc[0], acc[1], 0x4210);
                o[1] = __byte_perm(acc[1], acc[2], 0x5421);
                o[DST - 1] = __byte_perm(acc[2], acc[SRC - 1], 0x6542);
            }
            uint8_t* orow = out + (size_t)y * out_pitch;
            // output scanlines of any width (w * DST need not be a multiple of 4: 687 x 3 = 2061): the misalignment is the same for every
            // group of a scanline (groups are 4 * DST bytes apart), so the aligned words a thread's bytes straddle are assembled with
            // one funnel shift each, the word shared with the previous group from that lane's last word (shuffle); only the two ends
            // of a warp's stretch go out as single bytes
            const uint32_t prev_last = __shfl_up_sync(amask, o[DST - 1], 1);
            const uint32_t mis = (uint32_t)((uintptr_t)orow & 3u);
            if (full && (out_aligned || mis == 0u)) {
#pragma unroll
                for (int i = 0; i < DST; i++) *reinterpret_cast<uint32_t*>(orow + 4 * i) = o[i];
            } else if (full) {
                const uint32_t sh = 8u * (4u - mis);
                uint8_t* ab = orow - mis;                                        // 4-byte aligned
                if (lane > 0u) *reinterpret_cast<uint32_t*>(ab) = __funnelshift_r(prev_last, o[0], sh);
                else for (uint32_t b = 0; b < 4u - mis; b++) orow[b] = (uint8_t)(o[0] >> (8u * b));
#pragma unroll
                for (int i = 1; i < DST; i++) *reinterpret_cast<uint32_t*>(ab + 4 * i) = __funnelshift_r(o[i - 1], o[i], sh);
                if (!next_full) for (uint32_t b = 0; b < mis; b++) orow[4u * DST - mis + b] = (uint8_t)(o[DST - 1] >> (8u * (4u - mis + b)));
            } else {
#pragma unroll
                for (int i = 0; i < DST; i++) for (int b = 0; b < 4; b++) if ((uint32_t)(4 * i + b) < npx * DST) orow[4 * i + b] = (uint8_t)(o[i] >> (8 * b));
            }
        }
    }
}

__global__ void decode_status_kernel(DecodeParams p, uint32_t n)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) p.d_status[f] = p.state[f].status ? 1u /*FPNG_DECODE_NOT_FPNG*/ : 0u;
}

// Optional per-kernel timing of the decode pipeline (bench.py): events recorded between the launches of the most recent
// call while profiling is on; fpngb_decode_profile_read() returns the device times of prepare, scan, link, write, stored,
// unfilter (ms).
static bool g_dec_profile = false;
static cudaEvent_t g_dec_ev[8];
static bool g_dec_ev_ready = false, g_dec_ev_valid = false;
void decode_profile_enable(bool on)
{
    g_dec_profile = on;
    if (on && !g_dec_ev_ready) { for (int i = 0; i < 8; i++) cudaEventCreate(&g_dec_ev[i]); g_dec_ev_ready = true; }
}
int decode_profile_read(float* ms, int n)
{
    if (!g_dec_ev_valid) return 0;
    cudaEventSynchronize(g_dec_ev[6]);
    for (int i = 0; i < n && i < 6; i++) { ms[i] = 0.f; cudaEventElapsedTime(&ms[i], g_dec_ev[i], g_dec_ev[i + 1]); }
    return 1;
}
#define DEC_MARK(i) do { if (g_dec_profile && g_dec_ev_ready) cudaEventRecord(g_dec_ev[i], s); } while (0)

// FPNGB_DEC_STAGED=1 / fpngb_debug_decode_staged(1) put decode_write_staged_kernel in front of decode_write_kernel.  Off by default:
// MEASURED SLOWER on B200 (C2 write pass 3.06 vs 2.39 ms, C3 8.2 vs 6.3 ms).  The staging buffer limits an SM to 2 CTAs = 448 decode
// chains (the per-thread sink runs 1024), and the chain token length -> next window read -> next table entry is latency bound at that
// occupancy (ncu: issue active 54 %, 63 instructions per iteration against ~100 of the per-thread sink, both converged); byte-exact,
// covered by tests/test_decode_gpu.py::test_decode_write_paths_agree, numbers in profiles/README.md.
static bool g_dec_staged = getenv("FPNGB_DEC_STAGED") && atoi(getenv("FPNGB_DEC_STAGED")) != 0;
void decode_set_staged(bool on) { g_dec_staged = on; }

void launch_decode(const DecodeParams& p, uint32_t n, uint32_t desired, cudaStream_t s)
{
    DEC_MARK(0);
    decode_prepare_kernel<<<n, 32, 0, s>>>(p);
    DEC_MARK(1);
    const uint32_t sub_blocks = (p.subs_per_file + kDecThreads - 1) / kDecThreads;
    dim3 gsub(sub_blocks, n);
    constexpr size_t kDecSmem = (kFastWords + kWinSmemWords) * 4;
    FPNGB_SET_SMEM(decode_scan_kernel, kDecSmem);
    FPNGB_SET_SMEM(decode_write_kernel, kDecSmem);
    decode_scan_kernel<<<gsub, kDecThreads, kDecSmem, s>>>(p);
    DEC_MARK(2);
    decode_link_kernel<<<n, kLinkThreads, 0, s>>>(p);
    DEC_MARK(3);
    if (g_dec_staged) {
        FPNGB_SET_SMEM(decode_write_staged_kernel, kWsSmem);
        dim3 gws((p.subs_per_file + kWsThreads - 1) / kWsThreads, n);
        decode_write_staged_kernel<<<gws, kWsThreads, kWsSmem, s>>>(p);
    }
    decode_write_kernel<<<gsub, kDecThreads, kDecSmem, s>>>(p);
    DEC_MARK(4);
    dim3 gs((p.h + 7) / 8, n);
    decode_stored_kernel<<<gs, 256, 0, s>>>(p);
    DEC_MARK(5);
    const uint32_t groups = (p.w + 3) / 4;
    dim3 gu((groups + 127) / 128, n);
    if (p.chans == 3 && desired == 3) unfilter_kernel<3, 3><<<gu, 128, 0, s>>>(p);
    else if (p.chans == 3) unfilter_kernel<3, 4><<<gu, 128, 0, s>>>(p);
    else if (desired == 3) unfilter_kernel<4, 3><<<gu, 128, 0, s>>>(p);
    else unfilter_kernel<4, 4><<<gu, 128, 0, s>>>(p);
    decode_status_kernel<<<(n + 127) / 128, 128, 0, s>>>(p, n);
    DEC_MARK(6);
    if (g_dec_profile && g_dec_ev_ready) g_dec_ev_valid = true;
}

}  // namespace fpngb
