// oracle/ref_train_shim.cpp -- TEST INFRASTRUCTURE ONLY.  The unmodified reference compiled with its own
// FPNG_TRAIN_HUFFMAN_TABLES=1 switch (src/fpng.h:8-11, 114-120), exposing the table trainer for parity tests of
// fpngb_train_accumulate_device / fpngb_create_dynamic_block_prefix.  Built into oracle/_ref/libfpng_ref_train.so.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "fpng.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void reft_init(void) { fpng::fpng_init(); }
REF_API void reft_reset_counts(void) { memset(fpng::g_huff_counts, 0, sizeof(fpng::g_huff_counts)); }
REF_API void reft_get_counts(uint64_t* out288) { memcpy(out288, fpng::g_huff_counts, sizeof(fpng::g_huff_counts)); }
REF_API size_t reft_encode(const void* img, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags)
{
    std::vector<uint8_t> buf;
    if (!fpng::fpng_encode_image_to_memory(img, w, h, chans, buf, flags)) return 0;
    return buf.size();
}
REF_API int reft_create_prefix(const uint64_t* counts288, uint32_t chans, uint8_t* prefix, size_t cap, size_t* len,
                               uint64_t* bit_buf, int* bit_buf_size, uint32_t* codes288, uint8_t* sizes288)
{
    uint64_t freq[288];
    memcpy(freq, counts288, sizeof freq);
    std::vector<uint8_t> p;
    *bit_buf = 0; *bit_buf_size = 0;
    if (!fpng::create_dynamic_block_prefix(freq, chans, p, *bit_buf, *bit_buf_size, codes288, sizes288)) return 0;
    if (p.size() > cap) return 0;
    memcpy(prefix, p.data(), p.size());
    *len = p.size();
    return 1;
}
