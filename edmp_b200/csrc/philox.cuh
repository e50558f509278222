// Counter-based noise for the sampler's throughput mode (the parity tests replay recorded numpy draws instead).
#pragma once
#include <cstdint>

namespace edmp {

// ---- Philox4x32-10 (counter based, one normal per element and step) ---------------------------
__device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ double philox_normal(uint64_t seed, uint32_t step, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), step, 0x45444D50u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // Box-Muller on two 32-bit uniforms in (0, 1]
  const float u1 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;
  const float u2 = ((float)c[1] + 1.0f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  return (double)(r * cospif(2.0f * u2));
}

}  // namespace edmp
