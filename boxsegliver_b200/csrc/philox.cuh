// Counter-based random bits shared by the dropout (guide.cu) and input-noise (augment.cu) kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsl {
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator TF's random ops are built on. The stream here is
// keyed by (seed, offset) from the descriptor; element i uses counter i / 4, lane i % 4.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// TF's Uint32ToFloat: 23 random mantissa bits -> [1, 2) - 1 = [0, 1).
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long offset, unsigned long long idx) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)(idx >> 2), (unsigned)(idx >> 34), (unsigned)offset,
                                           (unsigned)(offset >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const unsigned lane = (unsigned)(idx & 3);
  const unsigned bits = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return __uint_as_float((bits & 0x7fffffu) | 0x3f800000u) - 1.0f;
}
}  // namespace bsl
