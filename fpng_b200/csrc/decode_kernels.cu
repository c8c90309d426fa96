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
    if (lane == 0) {
        // serial code assignment (ascending symbol order), parallel replication is not worth it at <= 288 symbols
        for (uint32_t i = 0; i < nsyms; i++) {
            const uint32_t l = sizes[i];
            if (!l) continue;
            const uint32_t code = __brev(s_next[l]++) >> (32 - l);
            if (l <= lut_bits) for (uint32_t c = code; c < lut_size; c += 1u << l) lut[c] = (uint16_t)(i | (l << 9));
        }
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
    const uint32_t f = blockIdx.x, lane = threadIdx.x;
    const FileDesc fd = p.files[f];
    DecodeState* st = p.state + f;
    const uint8_t* z = p.d_files + (size_t)f * p.file_stride + fd.idat_ofs + 8;
    BitSrc src{z, fd.file_size - (fd.idat_ofs + 8)};
    uint16_t* lut = p.luts + (size_t)f * 4096;

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
    if (!build_lut_warp(s_sizes, 288, lut, 12, lane, s_next)) { if (lane == 0) st->status = 1; }
}

// ------------------------------------------------------------------------------------------------
// D2: token decode
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kPosEnd = ~0ull;            // subsequence saw the end-of-block code
constexpr unsigned long long kPosErr = ~0ull - 1;        // subsequence hit an invalid code

struct SubResult {
    unsigned long long exit;     // first token boundary at/after the subsequence end, or kPosEnd / kPosErr
    unsigned long long eob_end;  // bit position right after the EOB code (valid when exit == kPosEnd)
    uint32_t n_out;              // filtered-stream bytes produced by tokens that start inside the subsequence
    uint32_t nlit;               // min(literal bytes produced, 4)
    uint32_t lits;               // the last <= 4 literal bytes, most recent in the top byte
};

// Decode tokens from `start` while the token's first bit is below `end`.  kWrite = false: bookkeeping only.
template <bool kWrite>
__device__ __forceinline__ SubResult decode_subsequence(const BitSrc& src, const uint16_t* __restrict__ s_lut, unsigned long long start,
                                                        unsigned long long end, uint32_t chans,
                                                        // write mode:
                                                        uint8_t* __restrict__ delta, uint32_t pitch, uint32_t bpl, uint32_t h,
                                                        unsigned long long out_pos, uint32_t tail, uint32_t* err)
{
    SubResult r; r.exit = start; r.eob_end = 0; r.n_out = 0; r.nlit = 0; r.lits = 0;
    if (start >= end) return r;                       // also covers kPosEnd / kPosErr pass-through
    BitCursor bc; bc.seek(src, start);
    uint32_t row = 0, col = 0;                        // position in the filtered stream (col 0 = filter byte)
    uint32_t lits = tail;                             // write mode: rolling window of the last 4 literal bytes
    if (kWrite) { row = (uint32_t)(out_pos / (bpl + 1ull)); col = (uint32_t)(out_pos % (bpl + 1ull)); }
    while (bc.pos < end) {
        bc.refill(src);
        const uint32_t e = s_lut[bc.peek(12)], l = e >> 9, s = e & 511;
        if (!l) { r.exit = kPosErr; return r; }
        bc.skip(l);
        if (s < 256) {
            r.n_out++;
            if (!kWrite) { r.lits = (r.lits >> 8) | (s << 24); r.nlit = min(r.nlit + 1u, 4u); }
            else {
                if (row >= h) { *err = 1; }
                else if (col == 0) { if (s != (row ? 2u : 0u)) *err = 1; }          // fpng.cpp:2264, 2642
                else delta[(size_t)row * pitch + (col - 1)] = (uint8_t)s;
                lits = (lits >> 8) | (s << 24);
                if (++col > bpl) { col = 0; row++; }
            }
        } else if (s == 256) {
            r.exit = kPosEnd; r.eob_end = bc.pos;
            return r;
        } else {
            if (s > 285) { r.exit = kPosErr; return r; }
            const uint32_t xb = c_len_xbits[s - 257];
            const uint32_t run = c_len_base[s - 257] + bc.peek(xb);
            bc.skip(xb + 1);                                                      // extra bits + the 1-bit distance code (fpng.cpp:2300)
            r.n_out += run;
            if (kWrite) {
                // run of the previous delta pixel: must start on a pixel boundary after at least one pixel of the row,
                // be a whole number of pixels and stay inside the row (fpng.cpp:2302-2315, 2681-2691, 2727)
                const bool bad = row >= h || col < 1 + chans || ((col - 1) % chans) != 0 || (run % chans) != 0 || (col - 1) + run > bpl;
                if (bad) { *err = 1; col += run; while (col > bpl) { col -= bpl + 1; row++; } }
                else {
                    uint8_t* d = delta + (size_t)row * pitch + (col - 1);
                    const uint32_t px = chans == 4 ? lits : (lits >> 8);          // last `chans` literals, oldest in the low byte
                    if (chans == 4) for (uint32_t i = 0; i < run; i += 4) { d[i] = (uint8_t)px; d[i + 1] = (uint8_t)(px >> 8); d[i + 2] = (uint8_t)(px >> 16); d[i + 3] = (uint8_t)(px >> 24); }
                    else for (uint32_t i = 0; i < run; i += 3) { d[i] = (uint8_t)px; d[i + 1] = (uint8_t)(px >> 8); d[i + 2] = (uint8_t)(px >> 16); }
                    col += run;
                    if (col > bpl) { col = 0; row++; }
                }
            }
        }
    }
    r.exit = bc.pos;
    return r;
}

// "keep the last 4 literal bytes" monoid: a then b
__device__ __forceinline__ void lit_combine(uint32_t& n, uint32_t& v, uint32_t nb, uint32_t vb)
{
    // v holds the last n literals with the most recent in the top byte
    if (nb >= 4) { n = 4; v = vb; return; }
    if (nb == 0) return;
    v = (v >> (8 * nb)) | vb;      // vb's valid bytes occupy its top nb bytes, low bytes are zero by construction
    n = min(n + nb, 4u);
}

__global__ void __launch_bounds__(kDecThreads) decode_tokens_kernel(DecodeParams p)
{
    __shared__ uint16_t s_lut[4096];
    __shared__ unsigned long long s_exit[kDecThreads];
    __shared__ unsigned long long s_scan[kDecThreads / 32];
    __shared__ uint32_t s_litn[kDecThreads / 32], s_litv[kDecThreads / 32];
    __shared__ unsigned long long s_carry_start, s_carry_out, s_eob_end;
    __shared__ uint32_t s_carry_litn, s_carry_litv, s_done, s_err;

    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DecodeState* st = p.state + f;
    if (st->status || st->stored) return;
    const FileDesc fd = p.files[f];
    const uint8_t* z = p.d_files + (size_t)f * p.file_stride + fd.idat_ofs + 8;
    BitSrc src{z, fd.file_size - (fd.idat_ofs + 8)};
    const uint32_t chans = p.chans, bpl = p.w * chans, h = p.h, pitch = p.delta_pitch;
    uint8_t* delta = p.delta + (size_t)f * pitch * h;
    const unsigned long long total_out = (unsigned long long)(bpl + 1) * h;

    for (uint32_t i = tid; i < 4096; i += blockDim.x) s_lut[i] = p.luts[(size_t)f * 4096 + i];
    if (tid == 0) { s_carry_start = st->token_start; s_carry_out = 0; s_carry_litn = 0; s_carry_litv = 0; s_done = 0; s_err = 0; s_eob_end = 0; }
    __syncthreads();

    const unsigned long long first_sub = s_carry_start / kSubBits;
    for (unsigned long long chunk = 0; ; chunk++) {
        const unsigned long long g = first_sub + chunk * kDecThreads + tid;       // absolute subsequence index
        const unsigned long long end = (g + 1) * kSubBits;
        unsigned long long start = tid == 0 ? s_carry_start : g * kSubBits;
        SubResult r = decode_subsequence<false>(src, s_lut, start, end, chans, nullptr, 0, 0, 0, 0, 0, nullptr);
        // iterate until every subsequence starts where its predecessor exits
        for (;;) {
            s_exit[tid] = r.exit;
            __syncthreads();
            const unsigned long long want = tid == 0 ? s_carry_start : s_exit[tid - 1];
            int changed = 0;
            if (want != start) {
                start = want;
                if (want == kPosEnd || want == kPosErr) { r.exit = want; r.n_out = 0; r.nlit = 0; r.lits = 0; r.eob_end = 0; }
                else r = decode_subsequence<false>(src, s_lut, start, end, chans, nullptr, 0, 0, 0, 0, 0, nullptr);
                changed = 1;
            }
            if (!__syncthreads_or(changed)) break;
        }
        // block-exclusive scans: output byte offsets and the literal window entering each subsequence
        unsigned long long incl = r.n_out;
        uint32_t wn = r.nlit, wv = r.lits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(kFullMask, incl, o);
            const uint32_t un = __shfl_up_sync(kFullMask, wn, o), uv = __shfl_up_sync(kFullMask, wv, o);
            if (lane >= (uint32_t)o) { incl += up; uint32_t tn = un, tv = uv; lit_combine(tn, tv, wn, wv); wn = tn; wv = tv; }
        }
        if (lane == 31) { s_scan[warp] = incl; s_litn[warp] = wn; s_litv[warp] = wv; }
        // exclusive values inside the warp
        unsigned long long ex = __shfl_up_sync(kFullMask, incl, 1);
        uint32_t en = __shfl_up_sync(kFullMask, wn, 1), ev = __shfl_up_sync(kFullMask, wv, 1);
        if (lane == 0) { ex = 0; en = 0; ev = 0; }
        __syncthreads();
        unsigned long long wbase = 0, chunk_total = 0;
        uint32_t bn = s_carry_litn, bv = s_carry_litv, cn = bn, cv = bv;
        for (uint32_t i = 0; i < kDecThreads / 32; i++) {
            if (i < warp) { wbase += s_scan[i]; lit_combine(bn, bv, s_litn[i], s_litv[i]); }
            chunk_total += s_scan[i];
            lit_combine(cn, cv, s_litn[i], s_litv[i]);
        }
        lit_combine(bn, bv, en, ev);                                             // window entering this subsequence
        const unsigned long long out_pos = s_carry_out + wbase + ex;

        // write pass
        uint32_t err = 0;
        if (r.exit == kPosErr) err = 1;
        if (r.n_out || r.exit == kPosEnd) {
            if (out_pos + r.n_out > total_out) err = 1;
            else decode_subsequence<true>(src, s_lut, start, end, chans, delta, pitch, bpl, h, out_pos, bv, &err);
        }
        if (r.exit == kPosEnd && start != kPosEnd) { s_eob_end = r.eob_end; s_done = 1; }
        if (err) s_err = 1;
        __syncthreads();
        if (tid == 0) {
            s_carry_start = s_exit[kDecThreads - 1];
            s_carry_out += chunk_total; s_carry_litn = cn; s_carry_litv = cv;
            if (s_carry_start == kPosErr) s_err = 1;
            if (s_carry_out > total_out) s_err = 1;
        }
        __syncthreads();
        if (s_done || s_err) break;
    }
    if (tid == 0) {
        // EOB right after the last row, then pad to a byte, then exactly the 4 Adler bytes (fpng.cpp:2559-2584)
        const unsigned long long used = (s_eob_end + 7) >> 3;
        const bool ok = !s_err && s_done && s_carry_out == total_out && used + 4 == fd.idat_len;
        st->status = ok ? 0u : 1u;
        st->out_bytes = s_carry_out; st->end_byte = used;
    }
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
    if (x0 >= w) return;
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
    for (uint32_t y = 0; y < h; y++) {
        uint32_t d[SRC];
        const uint8_t* dr = delta + (size_t)y * pitch;
        if (full) {
#pragma unroll
            for (int i = 0; i < SRC; i++) d[i] = *reinterpret_cast<const uint32_t*>(dr + 4 * i);   // pitch % 16 == 0 and x0*SRC % 4 == 0
        } else {
#pragma unroll
            for (int i = 0; i < SRC; i++) { d[i] = 0; for (int b = 0; b < 4; b++) if ((uint32_t)(4 * i + b) < npx * SRC) d[i] |= (uint32_t)dr[4 * i + b] << (8 * b); }
        }
#pragma unroll
        for (int i = 0; i < SRC; i++) acc[i] = summing ? vadd4(acc[i], d[i]) : d[i];
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
        if (full && out_aligned) {
#pragma unroll
            for (int i = 0; i < DST; i++) *reinterpret_cast<uint32_t*>(orow + 4 * i) = o[i];
        } else {
#pragma unroll
            for (int i = 0; i < DST; i++) for (int b = 0; b < 4; b++) if ((uint32_t)(4 * i + b) < npx * DST) orow[4 * i + b] = (uint8_t)(o[i] >> (8 * b));
        }
    }
}

__global__ void decode_status_kernel(DecodeParams p, uint32_t n)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) p.d_status[f] = p.state[f].status ? 1u /*FPNG_DECODE_NOT_FPNG*/ : 0u;
}

void launch_decode(const DecodeParams& p, uint32_t n, uint32_t desired, cudaStream_t s)
{
    decode_prepare_kernel<<<n, 32, 0, s>>>(p);
    decode_tokens_kernel<<<n, kDecThreads, 0, s>>>(p);
    dim3 gs((p.h + 7) / 8, n);
    decode_stored_kernel<<<gs, 256, 0, s>>>(p);
    const uint32_t groups = (p.w + 3) / 4;
    dim3 gu((groups + 127) / 128, n);
    if (p.chans == 3 && desired == 3) unfilter_kernel<3, 3><<<gu, 128, 0, s>>>(p);
    else if (p.chans == 3) unfilter_kernel<3, 4><<<gu, 128, 0, s>>>(p);
    else if (desired == 3) unfilter_kernel<4, 3><<<gu, 128, 0, s>>>(p);
    else unfilter_kernel<4, 4><<<gu, 128, 0, s>>>(p);
    decode_status_kernel<<<(n + 127) / 128, 128, 0, s>>>(p, n);
}

}  // namespace fpngb
