#include "../../include/fpng_b200.h"
extern "C" {
int fpngb_decode_host(const void*, uint32_t, void*, size_t, uint32_t*, uint32_t*, uint32_t*, uint32_t) { return FPNGB_DECODE_INVALID_ARG; }
int fpngb_decode_batch_device(const void*, size_t, const uint32_t*, const uint32_t*, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, void*, size_t, uint32_t*, void*) { return FPNGB_ERR_INTERNAL; }
}
