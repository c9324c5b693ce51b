// Counter-based Philox4x32-10 streams (Salmon et al., SC'11; Random123 reference constants).
// Replaces the reference's per-thread XORWOW state array (gPET_kernals.h:5, initialize.cu:256-272):
// no RNG state lives in HBM; a draw is a pure function of (seed, history index, photon, stage, block#).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpet {

struct Philox {
    uint32_t k0, k1;           // key = run seed
    uint32_t c0, c1, c2, c3;   // counter: (index lo, index hi, stream id, block number)

    __device__ __forceinline__ Philox(uint64_t seed, uint64_t index, uint32_t stream)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)index), c1((uint32_t)(index >> 32)),
          c2(stream), c3(0u) {}

    // next block of 4 x 32 random bits; advances the block number
    __device__ __forceinline__ uint4 next() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        c3++;
        return make_uint4(x0, x1, x2, x3);
    }
};

// (0,1] like curand_uniform: x * 2^-32 + 2^-33
__device__ __forceinline__ float u01(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
// (0,1) with 53 random bits
__device__ __forceinline__ double u01d(uint32_t a, uint32_t b) {
    uint64_t k = ((uint64_t)a << 21) | (uint64_t)(b >> 11);
    return (double)k * 1.1102230246251565e-16 + 5.551115123125783e-17;
}

}  // namespace gpet
