// Philox4x32-10 counter-based generator (Salmon et al., SC'11), hand-written for the passage kernels.
// Counter = (fish index, cell id lo, cell id hi, block number), key = 64-bit run seed, so every draw of every fish
// is a pure function of (seed, scenario/species/iteration/day cell id, fish, slot) and independent of how cells
// are partitioned over GPUs.
#pragma once
#include <stdint.h>

namespace stryke {

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  __host__ __device__ static __forceinline__ void round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                        uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }

  // 10 rounds; key schedule k += (W0, W1) per round (warp-uniform, so it lives in uniform registers / immediates).
  __host__ __device__ static __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c0, c1, c2, c3, k0, k1);
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

}  // namespace stryke
