// Counter-based random numbers shared by the device-side draws (throughput mode; the CPU generators of the reference
// cannot be reproduced on the device).
#pragma once
#include <stdint.h>

namespace pcvae {

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): counter-based, no state to initialise.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }          // [0, 1)
__device__ __forceinline__ float u01_open(uint32_t r) { return ((float)(r >> 8) + 1.0f) * 5.9604644775390625e-08f; }   // (0, 1]

// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float r = sqrtf(-2.0f * logf(u01_open(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}

}  // namespace pcvae
