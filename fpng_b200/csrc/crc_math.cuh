// fpng_b200/csrc/crc_math.cuh -- GF(2) arithmetic modulo the CRC-32 polynomial, shared by the checksum kernels and the
// single-pass encoder.  Values are in the reflected representation of the IEEE 802.3 CRC (fpng.cpp:195-249): bit 31 is
// the coefficient of x^0, so "1" is 0x80000000 and "x" is 0x40000000.  Advancing a CRC register over n zero bytes is the
// multiplication by x^(8n); NVIDIA GPUs have no carry-less multiply (the reference folds with PCLMULQDQ, fpng.cpp:255-281).
#pragma once
#include <stdint.h>

namespace fpngb {

constexpr uint32_t kCrcPoly = 0xEDB88320u;          // reflected IEEE 802.3 (fpng.cpp:195-249)
constexpr uint32_t kCrcOne = 0x80000000u;           // the polynomial "1" in reflected form
constexpr uint32_t kCrcXInv = 0xDB710641u;          // x^-1 mod P = (P - 1) / x

__host__ __device__ inline uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        p ^= (a & (0x80000000u >> i)) ? b : 0u;
        b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
    }
    return p;
}

__device__ inline uint32_t gf2_pow(const uint32_t* tab, unsigned long long e)
{
    uint32_t r = kCrcOne;
    for (int k = 0; e; k++, e >>= 1)
        if (e & 1ull) r = gf2_mulmod(r, tab[k]);
    return r;
}

}  // namespace fpngb
