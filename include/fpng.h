// include/fpng.h -- C++ drop-in surface of the B200-native fpng hot path.
//
// Source-compatible with the reference's public header (richgel999/fpng src/fpng.h:13-111): same namespace, function
// names, argument order/meaning, flag and status values, so an application (or fpng_test-style harness) that includes
// "fpng.h" and links libfpng_b200.so instead of compiling fpng.cpp keeps working.  Every function is a thin wrapper
// over the C ABI in fpng_b200.h; the work runs in CUDA kernels on the device selected by fpng_init().
// There is no CPU implementation behind these calls.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

#define FPNG_B200 1

namespace fpng
{
    // Library initialisation (reference: CPU feature detection, src/fpng.h:17).  Here: CUDA device/stream/table set-up.
    // Must be called once before any other function.  Uses the current CUDA device (FPNG_B200_DEVICE env overrides).
    void fpng_init();

    // Reference: "true if the CPU supports SSE 4.1" (src/fpng.h:23).  This build has no CPU SIMD path: always false.
    bool fpng_cpu_supports_sse41();

    const uint32_t FPNG_CRC32_INIT = 0;
    uint32_t fpng_crc32(const void* pData, size_t size, uint32_t prev_crc32 = FPNG_CRC32_INIT);   // src/fpng.h:27

    const uint32_t FPNG_ADLER32_INIT = 1;
    uint32_t fpng_adler32(const void* pData, size_t size, uint32_t adler = FPNG_ADLER32_INIT);     // src/fpng.h:31

    enum
    {
        FPNG_ENCODE_SLOWER = 1,        // per-image Huffman tables (2-pass), src/fpng.h:38
        FPNG_FORCE_UNCOMPRESSED = 2,   // stored Deflate blocks only, src/fpng.h:41
    };

    // src/fpng.h:48.  pImage: w*h pixels, num_chans (3|4) bytes each, R first, pitch w*num_chans, HOST memory.
    bool fpng_encode_image_to_memory(const void* pImage, uint32_t w, uint32_t h, uint32_t num_chans,
                                     std::vector<uint8_t>& out_buf, uint32_t flags = 0);

    // src/fpng.h:52
    bool fpng_encode_image_to_file(const char* pFilename, const void* pImage, uint32_t w, uint32_t h, uint32_t num_chans,
                                   uint32_t flags = 0);

    enum   // src/fpng.h:57-77
    {
        FPNG_DECODE_SUCCESS = 0,
        FPNG_DECODE_NOT_FPNG,
        FPNG_DECODE_INVALID_ARG,
        FPNG_DECODE_FAILED_NOT_PNG,
        FPNG_DECODE_FAILED_HEADER_CRC32,
        FPNG_DECODE_FAILED_INVALID_DIMENSIONS,
        FPNG_DECODE_FAILED_DIMENSIONS_TOO_LARGE,
        FPNG_DECODE_FAILED_CHUNK_PARSING,
        FPNG_DECODE_FAILED_INVALID_IDAT,
        FPNG_DECODE_FILE_OPEN_FAILED,
        FPNG_DECODE_FILE_TOO_LARGE,
        FPNG_DECODE_FILE_READ_FAILED,
        FPNG_DECODE_FILE_SEEK_FAILED
    };

    // src/fpng.h:92
    int fpng_get_info(const void* pImage, uint32_t image_size, uint32_t& width, uint32_t& height, uint32_t& channels_in_file);

    // src/fpng.h:108.  Decodes files written by fpng (this library or the reference); anything else -> FPNG_DECODE_NOT_FPNG.
    int fpng_decode_memory(const void* pImage, uint32_t image_size, std::vector<uint8_t>& out, uint32_t& width, uint32_t& height,
                           uint32_t& channels_in_file, uint32_t desired_channels);

    // src/fpng.h:111
    int fpng_decode_file(const char* pFilename, std::vector<uint8_t>& out, uint32_t& width, uint32_t& height,
                         uint32_t& channels_in_file, uint32_t desired_channels);
}
