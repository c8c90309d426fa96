// fpng_b200/csrc/decode.cuh -- parameter blocks of the decode kernels.
#pragma once
#include "common.cuh"

namespace fpngb {

constexpr int kDecThreads = 256;            // subsequences per CTA in the scan / write kernels
constexpr int kLinkThreads = 1024;          // subsequences linked per step by the per-file chain kernel
constexpr uint32_t kSubBits = 1024;         // bits per subsequence (>> longest token: 12 + 5 + 1 bits)
constexpr uint32_t kPreRoll = 256;          // speculative lead-in before a subsequence (self-synchronisation distance)

// Per-subsequence record.  After decode_scan_kernel: start/exit/eob_end are absolute aligned-stream bit positions,
// n_out/nlit/lits describe the tokens starting in the subsequence.  decode_link_kernel rewrites it for the write
// kernel: exit := output byte offset, lits := the 4 literals preceding the subsequence, nlit := live flag.
struct SubInfo {
    unsigned long long start, exit, eob_end;
    uint32_t n_out, lits, nlit, pad_;
};

struct FileDesc {                           // filled by the host container walk (fpng.cpp:2930-3077)
    uint32_t file_size, idat_ofs, idat_len, pad_;
};

struct DecodeState {
    uint32_t status;                        // 0 ok, 1 -> FPNG_DECODE_NOT_FPNG
    uint32_t stored;                        // stored-block file
    unsigned long long token_start;         // bit offset (from the zlib stream start) of the first token
    unsigned long long out_bytes;           // filtered-stream bytes produced
    unsigned long long end_byte;            // bytes consumed up to and including the padded EOB
};

constexpr uint32_t kFastWords = 4096 + 64;   // 4096 fast entries, then the 256 literal code sizes (bytes)

struct DecodeParams {
    const uint8_t* d_files; size_t file_stride;
    const FileDesc* files;                  // [n] device
    DecodeState* state;                     // [n]
    uint32_t* luts;                         // [n][4096]  fused entries, see decode_kernels.cu
    uint32_t* fast;                         // [n][kFastWords]  multi-token entries + literal code sizes of the scan / write loops
    SubInfo* subs; uint32_t subs_per_file;  // [n][subs_per_file]
    uint8_t* delta; uint32_t delta_pitch;   // [n][h][pitch] filtered rows without the filter byte
    uint8_t* d_out; size_t out_stride;
    uint32_t* d_status;                     // [n] FPNG_DECODE_* codes
    uint32_t w, h, chans;
};

void launch_decode(const DecodeParams& p, uint32_t n, uint32_t desired, cudaStream_t s);
void decode_profile_enable(bool on);
int decode_profile_read(float* ms, int n);
unsigned long long decode_link_repairs(bool reset);
void decode_set_staged(bool on);            // write pass through the shared-memory staged kernel (default) or decode_write_kernel alone

}  // namespace fpngb
