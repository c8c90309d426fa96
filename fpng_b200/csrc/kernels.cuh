// fpng_b200/csrc/kernels.cuh -- kernel parameter blocks and launcher prototypes (host_api.cu <-> *_kernels.cu).
#pragma once
#include "common.cuh"

namespace fpngb {

constexpr int kRowsPerBlock = 8;                  // one warp per scanline, 8 scanlines per CTA
constexpr int kScanThreads = 32 * kRowsPerBlock;
constexpr int kPackThreads = 32 * kRowsPerBlock;
constexpr int kOffsetsThreads = 256;
constexpr int kStageWords = 272;                  // >= (31 + 12 + 128 * 66) / 32 + 1 words per 128-pixel step

struct ScanParams {
    const uint8_t* pixels; size_t image_stride;   // n images, tightly packed rows (pitch = w * chans)
    uint32_t w, h;
    const CodeBook* books; uint32_t book_stride;  // 0: one shared (1-pass) book, 1: per-image books (2-pass)
    uint32_t* row_bits;                           // [n*h] token bits per row (incl. the filter literal)
    uint2* row_adler;                             // [n*h] (S1, S2) mod 65521 of the filtered row
    ImageState* st;                               // [n]
    uint32_t* hist;                               // [n*288] (histogram mode)
    uint32_t merge_first_unit;                    // RGB 1-pass: filter literal and pixel 0 share a flush unit (fpng.cpp:1187-1203)
    uint2* lane_ofs; uint32_t lane_ofs_pitch;     // v2 kernels: [n*h][pitch] per 16-pixel group: .x bit offset inside its row, .y lane_info16()
    uint32_t lit1_rule;                           // RGBA 1-pass with a table under which fpng.cpp:1520-1528 can fire (generic kernels only)
};

struct OffsetsParams {
    const uint32_t* row_bits; unsigned long long* row_ofs;
    const CodeBook* books; uint32_t book_stride;
    ImageState* st;
    uint8_t* out; size_t out_stride; uint32_t* sizes;
    uint32_t w, h, chans, flags;
    uint8_t png_header[kPngHeaderSize];           // IDAT length patched per image on device
};

struct PackParams {
    const uint8_t* pixels; size_t image_stride;
    uint32_t w, h;
    const CodeBook* books; uint32_t book_stride;
    const unsigned long long* row_ofs;
    const uint32_t* row_bits;                     // v2 kernels: total token bits per row
    const uint2* lane_ofs; uint32_t lane_ofs_pitch;
    uint2* row_adler;                             // rewritten for stored images (raw bytes, filter 0)
    const ImageState* st;
    uint8_t* out; size_t out_stride;
    uint32_t lit1_rule;                           // see ScanParams
    uint32_t stored_only;                         // 16-pixel pack kernel: only write images that fell back to stored blocks (fused encoder ran before)
    // 16-pixel pack kernel: raw CRC-32 of each scanline's code words, computed while they sit in the staging buffer
    // (nullptr = off; the file-reading CRC kernel runs instead).  Tables: crc_stream_kernel.cu.
    uint32_t* row_crc;                            // [n*h]
    const uint32_t* crc_f128b;                    // [4][256] slice tables of the 128-byte advance
    const uint32_t* crc_lane_mul;                 // [32][8][16] nibble tables of x^(32 * (32 - l))
};

// combines the scanline CRCs of the pack kernel into the IDAT CRC (crc_stream_kernel.cu)
struct RowCrcParams {
    const uint32_t* row_crc; const unsigned long long* row_ofs; const uint32_t* row_bits;
    const CodeBook* books; uint32_t book_stride;
    ImageState* st; uint8_t* out; size_t out_stride;
    uint32_t h;
};

struct AdlerParams {
    const uint2* row_adler; ImageState* st;
    uint8_t* out; size_t out_stride;
    uint32_t w, h, chans;
};

struct CrcParams {
    uint8_t* out; size_t out_stride;
    ImageState* st;
    uint32_t max_tiles;                           // grid.x; CTAs beyond an image's length exit immediately
    uint32_t msg_start;                           // first byte of the CRC'd region (54 = "IDAT" for a PNG file)
    uint32_t init_xor;                            // initial register value (0xFFFFFFFF for a fresh CRC)
    uint32_t stored_only;                         // only images that fell back to stored blocks (the single-pass encoder computed the others' CRC inline)
};

struct HuffParams {
    const uint32_t* hist;                         // [n*288]
    CodeBook* books;                              // [n]
    uint32_t chans;
    uint32_t training;                            // 1: table training (fpng.cpp:909-988): symbol 256 keeps its own count in the scaling
};

// opt in to > 48 KiB of dynamic shared memory (per kernel instantiation, once per process)
#define FPNGB_SET_SMEM(kernel, bytes) do { static bool done_ = false; \
    if (!done_) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); done_ = true; } } while (0)

// third-generation single-pass encoder (encode_fused.cu)
bool fused_eligible(const void* base, size_t image_stride, uint32_t w, uint32_t h, uint32_t chans, uint32_t n);
size_t fused_desc_bytes(uint32_t n, uint32_t w, uint32_t h);
int launch_encode_fused(const uint8_t* pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags,
                        const CodeBook* books, uint32_t book_stride, uint2* row_adler, ImageState* st, void* desc_mem,
                        uint8_t* out, size_t out_stride, uint32_t* sizes, const uint8_t* png_header, uint32_t merge_first_unit, cudaStream_t s,
                        cudaEvent_t mid_event, bool inline_crc);
void launch_fused_crc(const void* desc_mem, uint32_t n, uint32_t w, uint32_t h, const CodeBook* books, uint32_t book_stride, const ImageState* st,
                      uint8_t* out, size_t out_stride, cudaStream_t s);
int fused_tables_init();
// second-generation IDAT CRC kernel (crc_stream_kernel.cu)
int crc_stream_tables_init();
void launch_crc_stream(const CrcParams& p, uint32_t n, size_t max_file_bytes, cudaStream_t s);
void launch_row_crc_combine(const RowCrcParams& p, uint32_t n, cudaStream_t s);
int crc_stream_table_ptrs(const uint32_t** f128b, const uint32_t** lane_mul);

void launch_scan(const ScanParams& p, uint32_t n, uint32_t chans, int mode, bool hist, cudaStream_t s);
bool walk16_eligible(const void* base, size_t image_stride, uint32_t w, uint32_t chans);
void launch_scan16(const ScanParams& p, uint32_t n, uint32_t chans, cudaStream_t s);
void launch_hist16(const ScanParams& p, uint32_t n, uint32_t chans, cudaStream_t s);
void launch_pack16(const PackParams& p, uint32_t n, uint32_t chans, cudaStream_t s);
void set_rows_per_warp16(uint32_t v);          // 0 = automatic (5 scanlines per warp on large batches, 1 otherwise)
void launch_offsets(const OffsetsParams& p, uint32_t n, cudaStream_t s);
void launch_pack(const PackParams& p, uint32_t n, uint32_t chans, int mode, cudaStream_t s);
void launch_adler_finalize(const AdlerParams& p, uint32_t n, cudaStream_t s);
void launch_crc(const CrcParams& p, uint32_t n, cudaStream_t s);
void launch_huffman_build(const HuffParams& p, uint32_t n, cudaStream_t s);
uint32_t crc_ctas_for(size_t max_file_bytes);
uint32_t adler_chunk_bytes();
void launch_adler_buffer(const uint8_t* d_buf, size_t n, uint2* d_partials, cudaStream_t s);
void launch_compact(const uint8_t* files, size_t stride, const uint32_t* sizes, uint32_t n, uint8_t* dst, size_t dst_cap,
                    unsigned long long* offsets, cudaStream_t s);
int  checksum_tables_init();                      // uploads CRC slice tables / x^(2^k) powers (idempotent)

// host helpers shared by host_api.cu and the tests
uint32_t host_crc32(const void* data, size_t n, uint32_t prev);

}  // namespace fpngb
