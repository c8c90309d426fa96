#include "../../include/fpng_b200.h"
extern "C" {
int fpngb_encode_batch_host(const void*, size_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, void*, size_t, uint32_t*) { return FPNGB_ERR_INTERNAL; }
int fpngb_decode_host(const void*, uint32_t, void*, size_t, uint32_t*, uint32_t*, uint32_t*, uint32_t) { return FPNGB_DECODE_INVALID_ARG; }
int fpngb_decode_batch_device(const void*, size_t, const uint32_t*, const uint32_t*, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, void*, size_t, uint32_t*, void*) { return FPNGB_ERR_INTERNAL; }
uint32_t fpngb_crc32(const void*, size_t, uint32_t) { return 0; }
uint32_t fpngb_adler32(const void*, size_t, uint32_t) { return 0; }
}
