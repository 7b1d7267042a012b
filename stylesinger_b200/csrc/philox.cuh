// Counter-based RNG for the production (non-parity) sampling mode: Philox4x32-10 keyed by
// (seed, stream id), counter = element index.  Parity tests inject explicit noise tensors instead
// (SURVEY.md §7 "RNG parity"), so this generator never has to match torch's.
#pragma once
#include <stdint.h>

namespace ssb {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t stream, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint32_t o[4];
  philox4(seed, stream, idx, o);
  return u32_to_unit(o[0]);
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint32_t o[4];
  philox4(seed, stream, idx, o);
  const float u1 = u32_to_unit(o[0]), u2 = u32_to_unit(o[1]);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace ssb
