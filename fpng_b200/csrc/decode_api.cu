// fpng_b200/csrc/decode_api.cu -- container walk (host) and decode entry points of the C ABI.
#include "../../include/fpng_b200.h"
#include "kernels.cuh"
#include "decode.cuh"
#include "runtime.h"
#include <string.h>
#include <vector>

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

namespace fpngb {

static Buffer g_dec_ws;        // FileDesc[n] | DecodeState[n] | luts | delta scratch
static Buffer g_dec_pin;       // pinned staging for the FileDesc upload / status download

// Enqueue the decode pipeline for n files already resident on the device.  Caller holds context().mu.
static int decode_batch_locked(const uint8_t* d_files, size_t file_stride, const FileDesc* h_files, uint32_t n, uint32_t w, uint32_t h,
                               uint32_t chans, uint32_t desired, uint8_t* d_out, size_t out_stride, uint32_t* d_status, cudaStream_t s)
{
    const uint32_t bpl = w * chans, pitch = (uint32_t)align_up(bpl, 16);
    size_t o = 0;
    const size_t o_files = o; o = align_up(o + (size_t)n * sizeof(FileDesc), 256);
    const size_t o_state = o; o = align_up(o + (size_t)n * sizeof(DecodeState), 256);
    const size_t o_luts = o; o = align_up(o + (size_t)n * 4096 * 4, 256);
    const size_t o_fast = o; o = align_up(o + (size_t)n * kFastWords * 4, 256);
    uint32_t max_idat = 0;
    for (uint32_t i = 0; i < n; i++) if (h_files[i].idat_len > max_idat) max_idat = h_files[i].idat_len;
    const uint32_t subs_per_file = (uint32_t)(((uint64_t)max_idat * 8 + 32) / kSubBits + 2);
    const size_t o_subs = o; o = align_up(o + (size_t)n * subs_per_file * sizeof(SubInfo), 256);
    const size_t o_delta = o; o = align_up(o + (size_t)n * pitch * h + 64, 256);
    int rc = g_dec_ws.reserve(o); if (rc) return rc;
    rc = context().ws_acquire(s); if (rc) return rc;
    g_dec_pin.pinned = true;
    rc = g_dec_pin.reserve((size_t)n * sizeof(FileDesc)); if (rc) return rc;
    // the pinned staging buffer may still be in flight from the previous call on another stream: wait for it
    static cudaEvent_t staged = nullptr;
    if (!staged) FPNGB_CUDA_OK(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
    else FPNGB_CUDA_OK(cudaEventSynchronize(staged));
    memcpy(g_dec_pin.p, h_files, (size_t)n * sizeof(FileDesc));
    uint8_t* b = (uint8_t*)g_dec_ws.p;
    FPNGB_CUDA_OK(cudaMemcpyAsync(b + o_files, g_dec_pin.p, (size_t)n * sizeof(FileDesc), cudaMemcpyHostToDevice, s));
    FPNGB_CUDA_OK(cudaEventRecord(staged, s));
    DecodeParams p{};
    p.d_files = d_files; p.file_stride = file_stride; p.files = (const FileDesc*)(b + o_files); p.state = (DecodeState*)(b + o_state);
    p.luts = (uint32_t*)(b + o_luts); p.fast = (uint32_t*)(b + o_fast); p.subs = (SubInfo*)(b + o_subs); p.subs_per_file = subs_per_file; p.delta = b + o_delta; p.delta_pitch = pitch; p.d_out = d_out; p.out_stride = out_stride;
    p.d_status = d_status; p.w = w; p.h = h; p.chans = chans;
    launch_decode(p, n, desired, s);
    count_launch(8);
    FPNGB_CUDA_OK(cudaGetLastError());
    return context().ws_release(s);
}

}  // namespace fpngb

using namespace fpngb;

extern "C" {

int fpngb_decode_batch_device(const void* d_files, size_t file_stride, const uint32_t* file_sizes, const uint32_t* idat_ofs,
                              const uint32_t* idat_len, uint32_t n, uint32_t w, uint32_t h, uint32_t chans_in_file,
                              uint32_t desired_chans, void* d_out, size_t out_stride, uint32_t* d_status, void* stream)
{
    Context& c = context();
    if (!c.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!d_files || !file_sizes || !idat_ofs || !idat_len || !d_out || !d_status || n == 0 || n > 65535) return FPNGB_ERR_INVALID_ARG;
    if ((chans_in_file != 3 && chans_in_file != 4) || (desired_chans != 3 && desired_chans != 4) || !w || !h) return FPNGB_ERR_INVALID_ARG;
    // the container limits of fpng_get_info (fpng.cpp:2966-2971): keeps w * chans and the scratch pitch inside 32 bits
    if (w > (1u << 24) || h > (1u << 24) || (uint64_t)w * h > (1ull << 30)) return FPNGB_ERR_INVALID_ARG;
    if ((uint64_t)w * h * desired_chans > 0xFFFFFFFFull || out_stride < (size_t)w * h * desired_chans) return FPNGB_ERR_BUFFER_TOO_SMALL;
    if (file_stride % 4 || (uintptr_t)d_files % 4) return FPNGB_ERR_ALIGNMENT;
    std::vector<FileDesc> fds(n);
    for (uint32_t i = 0; i < n; i++) {
        if ((uint64_t)idat_ofs[i] + 8 + idat_len[i] + 4 > file_sizes[i] || file_sizes[i] > file_stride) return FPNGB_ERR_INVALID_ARG;
        fds[i] = FileDesc{file_sizes[i], idat_ofs[i], idat_len[i], 0};
    }
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    return decode_batch_locked((const uint8_t*)d_files, file_stride, fds.data(), n, w, h, chans_in_file, desired_chans, (uint8_t*)d_out,
                               out_stride, d_status, (cudaStream_t)stream);
}

int fpngb_decode_host(const void* file, uint32_t size, void* out, size_t out_cap, uint32_t* w, uint32_t* h, uint32_t* chans, uint32_t desired)
{
    uint32_t ww = 0, hh = 0, cc = 0, idat_ofs = 0, idat_len = 0;
    if (w) *w = 0; if (h) *h = 0; if (chans) *chans = 0;
    if (!file || !size || !out || (desired != 3 && desired != 4)) return FPNGB_DECODE_INVALID_ARG;     // fpng.cpp:3092-3096
    const int st = container_info((const uint8_t*)file, size, &ww, &hh, &cc, &idat_ofs, &idat_len);
    if (w) *w = ww; if (h) *h = hh; if (chans) *chans = cc;
    if (st) return st;
    const uint64_t need = (uint64_t)ww * hh * desired;
    if (need > 0xFFFFFFFFull) return FPNGB_DECODE_FAILED_DIMENSIONS_TOO_LARGE;                          // fpng.cpp:3103-3105
    if (out_cap < need) return FPNGB_DECODE_INVALID_ARG;
    Context& c = context();
    if (!c.ready) return FPNGB_DECODE_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c.mu);
    if (cudaSetDevice(c.device) != cudaSuccess) return FPNGB_DECODE_INVALID_ARG;
    const size_t fstride = align_up((size_t)size + 16, 16);
    if (c.dev_in.reserve(fstride) || c.dev_out.reserve(align_up(need, 16) + 64)) return FPNGB_DECODE_INVALID_ARG;
    uint32_t* d_status = (uint32_t*)((uint8_t*)c.dev_out.p + align_up(need, 16));
    cudaStream_t s = c.stream;
    cudaMemcpyAsync(c.dev_in.p, file, size, cudaMemcpyHostToDevice, s);
    FileDesc fd{size, idat_ofs, idat_len, 0};
    if (decode_batch_locked((const uint8_t*)c.dev_in.p, fstride, &fd, 1, ww, hh, cc, desired, (uint8_t*)c.dev_out.p, align_up(need, 16), d_status, s))
        return FPNGB_DECODE_INVALID_ARG;
    uint32_t* h_status = (uint32_t*)c.pin_small.p;
    cudaMemcpyAsync(h_status, d_status, 4, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return FPNGB_DECODE_INVALID_ARG;
    if (*h_status) return FPNGB_DECODE_NOT_FPNG;                                                        // fpng.cpp:3131-3136
    cudaMemcpyAsync(out, c.dev_out.p, need, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return FPNGB_DECODE_INVALID_ARG;
    return FPNGB_DECODE_SUCCESS;
}

int fpngb_get_info(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans)
{
    uint32_t a, b, ww = 0, hh = 0, cc = 0;
    if (!file) { if (w) *w = 0; if (h) *h = 0; if (chans) *chans = 0; return FPNGB_DECODE_INVALID_ARG; }
    const int st = container_info((const uint8_t*)file, size, &ww, &hh, &cc, &a, &b);
    if (w) *w = ww; if (h) *h = hh; if (chans) *chans = cc;
    return st;
}

// Pipelined host-buffer batch decode: per chunk the files go H2D, the decode kernels run, the pixels come back D2H;
// three slots keep copies and kernels of neighbouring chunks overlapped.  Files that fail the container walk keep their
// status and are skipped; all decodable files must share width/height/channels.
int fpngb_decode_batch_host(const void* const* files, const uint32_t* sizes, uint32_t n, uint32_t desired, void* out, size_t out_stride,
                            uint32_t* w_out, uint32_t* h_out, uint32_t* chans_out, int* status)
{
    Context& c = context();
    if (!c.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!files || !sizes || !out || !status || n == 0 || (desired != 3 && desired != 4)) return FPNGB_ERR_INVALID_ARG;
    std::vector<FileDesc> fds(n);
    uint32_t W = 0, H = 0, C = 0, max_size = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t ww = 0, hh = 0, cc = 0, io = 0, il = 0;
        status[i] = (!files[i] || !sizes[i]) ? FPNGB_DECODE_INVALID_ARG : container_info((const uint8_t*)files[i], sizes[i], &ww, &hh, &cc, &io, &il);
        fds[i] = FileDesc{sizes[i], io, il, 0};
        if (status[i]) continue;
        if (!W) { W = ww; H = hh; C = cc; }
        else if (ww != W || hh != H || cc != C) return FPNGB_ERR_INVALID_ARG;
        if (sizes[i] > max_size) max_size = sizes[i];
    }
    if (w_out) *w_out = W; if (h_out) *h_out = H; if (chans_out) *chans_out = C;
    if (!W) return FPNGB_OK;                                   // nothing decodable: statuses say why
    const uint64_t need = (uint64_t)W * H * desired;
    if (need > 0xFFFFFFFFull) { for (uint32_t i = 0; i < n; i++) if (!status[i]) status[i] = FPNGB_DECODE_FAILED_DIMENSIONS_TOO_LARGE; return FPNGB_OK; }
    if (out_stride < need) return FPNGB_ERR_BUFFER_TOO_SMALL;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));

    constexpr int kSlots = 3;
    const size_t fstride = align_up((size_t)max_size + 16, 16), pstride = align_up((size_t)need, 16);
    uint32_t per_chunk = (uint32_t)(((size_t)64 << 20) / need); if (per_chunk < 1) per_chunk = 1; if (per_chunk > n) per_chunk = n;
    if (per_chunk > 65535u) per_chunk = 65535u;                     // grid.y limit of the batch kernels (tiny images)
    const size_t slot_in = per_chunk * fstride, slot_out = per_chunk * pstride + align_up(per_chunk * 4, 256);
    int rc = c.dev_in.reserve(kSlots * slot_in); if (rc) return rc;
    rc = c.dev_out.reserve(kSlots * slot_out); if (rc) return rc;
    rc = c.pin_small.reserve(kSlots * align_up(per_chunk * 4, 256)); if (rc) return rc;
    if (!c.copy_in) { FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking)); FPNGB_CUDA_OK(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking)); }
    EventSet<3 * kSlots> evs;                                       // destroyed on every exit path
    if (!evs.ok) return FPNGB_ERR_INTERNAL;
    cudaEvent_t* ev_in = evs.ev; cudaEvent_t* ev_done = evs.ev + kSlots; cudaEvent_t* ev_out = evs.ev + 2 * kSlots;
    // decodable files, in order
    std::vector<uint32_t> idx; idx.reserve(n);
    for (uint32_t i = 0; i < n; i++) if (!status[i]) idx.push_back(i);
    const uint32_t m = (uint32_t)idx.size(), nchunks = (m + per_chunk - 1) / per_chunk;
    int result = FPNGB_OK;
    std::vector<FileDesc> cf(per_chunk);
    for (uint32_t k = 0; k < nchunks && result == FPNGB_OK; k++) {
        const int sl = k % kSlots;
        const uint32_t first = k * per_chunk, cnt = (first + per_chunk <= m) ? per_chunk : m - first;
        if (k >= kSlots) {   // the slot's previous pixels and statuses have left the device
            if (cudaEventSynchronize(ev_out[sl]) != cudaSuccess) { result = FPNGB_ERR_INTERNAL; break; }
            const uint32_t pf = (k - kSlots) * per_chunk, pc = (pf + per_chunk <= m) ? per_chunk : m - pf;
            const uint32_t* hs = (const uint32_t*)((uint8_t*)c.pin_small.p + sl * align_up(per_chunk * 4, 256));
            for (uint32_t j = 0; j < pc; j++) status[idx[pf + j]] = hs[j] ? FPNGB_DECODE_NOT_FPNG : FPNGB_DECODE_SUCCESS;
        }
        uint8_t* din = (uint8_t*)c.dev_in.p + sl * slot_in;
        uint8_t* dout = (uint8_t*)c.dev_out.p + sl * slot_out;
        uint32_t* dst = (uint32_t*)(dout + per_chunk * pstride);
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t i = idx[first + j];
            cudaMemcpyAsync(din + j * fstride, files[i], sizes[i], cudaMemcpyHostToDevice, c.copy_in);
            cf[j] = fds[i];
        }
        cudaEventRecord(ev_in[sl], c.copy_in);
        cudaStreamWaitEvent(c.stream, ev_in[sl], 0);
        rc = decode_batch_locked(din, fstride, cf.data(), cnt, W, H, C, desired, dout, pstride, dst, c.stream);
        if (rc) { result = rc; break; }
        cudaEventRecord(ev_done[sl], c.stream);
        cudaStreamWaitEvent(c.copy_out, ev_done[sl], 0);
        cudaMemcpyAsync((uint8_t*)c.pin_small.p + sl * align_up(per_chunk * 4, 256), dst, cnt * 4, cudaMemcpyDeviceToHost, c.copy_out);
        for (uint32_t j = 0; j < cnt; j++)
            cudaMemcpyAsync((uint8_t*)out + (size_t)idx[first + j] * out_stride, dout + j * pstride, need, cudaMemcpyDeviceToHost, c.copy_out);
        cudaEventRecord(ev_out[sl], c.copy_out);
    }
    cudaStreamSynchronize(c.copy_in); cudaStreamSynchronize(c.stream); cudaStreamSynchronize(c.copy_out);
    if (result == FPNGB_OK) {
        for (uint32_t k = (nchunks > kSlots ? nchunks - kSlots : 0); k < nchunks; k++) {
            const int sl = k % kSlots;
            const uint32_t pf = k * per_chunk, pc = (pf + per_chunk <= m) ? per_chunk : m - pf;
            const uint32_t* hs = (const uint32_t*)((uint8_t*)c.pin_small.p + sl * align_up(per_chunk * 4, 256));
            for (uint32_t j = 0; j < pc; j++) status[idx[pf + j]] = hs[j] ? FPNGB_DECODE_NOT_FPNG : FPNGB_DECODE_SUCCESS;
        }
        if (cudaGetLastError() != cudaSuccess) result = FPNGB_ERR_INTERNAL;
    }
    return result;
}

// per-kernel device times (ms) of the most recent decode call made while profiling was on: prepare, scan, link, write, stored, unfilter
FPNGB_API int fpngb_decode_profile_enable(int on) { decode_profile_enable(on != 0); return 0; }
FPNGB_API int fpngb_decode_profile_read(float* ms, int n) { return decode_profile_read(ms, n); }
// test hook: number of subsequences the link pass had to decode again since the last reset (synchronises the device)
FPNGB_API void fpngb_debug_decode_staged(int on) { decode_set_staged(on != 0); }   // tests / A/B: 0 = decode_write_kernel alone
FPNGB_API unsigned long long fpngb_debug_decode_repairs(int reset) { cudaDeviceSynchronize(); return decode_link_repairs(reset != 0); }

int fpngb_get_info_ex(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans, uint32_t* idat_ofs, uint32_t* idat_len)
{
    if (!file || !w || !h || !chans || !idat_ofs || !idat_len) return FPNGB_DECODE_INVALID_ARG;
    return container_info((const uint8_t*)file, size, w, h, chans, idat_ofs, idat_len);
}

}  // extern "C"
