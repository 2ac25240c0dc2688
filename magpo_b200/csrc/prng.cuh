// Device-side jax.random (threefry2x32, partitionable scheme) — SURVEY.md Appendix A1-A4.
#pragma once
#include <stdint.h>

namespace magpo {

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// Threefry-2x32, 20 rounds (Appendix A1).
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                                                      uint32_t& o0, uint32_t& o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define TF_ROUND(r) \
  x0 += x1;         \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  x0 += k0; x1 += k1;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
#undef TF_ROUND
  o0 = x0; o1 = x1;
}

// split(key, n)[i]  (Appendix A2, partitionable): threefry(key, (hi32(i), lo32(i))).
__host__ __device__ __forceinline__ void prng_split_i(uint32_t k0, uint32_t k1, uint32_t i, uint32_t& o0, uint32_t& o1) {
  threefry2x32(k0, k1, 0u, i, o0, o1);
}
// random_bits(key, 32, shape)[i] = out0 ^ out1.
__host__ __device__ __forceinline__ uint32_t prng_bits_i(uint32_t k0, uint32_t k1, uint64_t i) {
  uint32_t o0, o1;
  threefry2x32(k0, k1, (uint32_t)(i >> 32), (uint32_t)i, o0, o1);
  return o0 ^ o1;
}
// uniform(key, minval=tiny, maxval=1)[i] then gumbel = -log(-log(u))  (Appendix A3).
__device__ __forceinline__ float prng_uniform_from_bits(uint32_t bits) {
  float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  return fmaxf(1.17549435e-38f, f + 1.17549435e-38f);  // f*(1-tiny)+tiny; max(tiny, .)
}
__device__ __forceinline__ float prng_gumbel_from_bits(uint32_t bits) {
  return -logf(-logf(prng_uniform_from_bits(bits)));
}
// randint offset from two 32-bit draws (Appendix A4).
__host__ __device__ __forceinline__ int32_t prng_randint_from_bits(uint32_t hi, uint32_t lo, int32_t minval, int32_t maxval) {
  uint32_t span = maxval > minval ? (uint32_t)(maxval - minval) : 1u;
  uint32_t m = 65536u % span;
  uint32_t mult = (m * m) % span;
  uint32_t off = ((hi % span) * mult + (lo % span)) % span;
  return minval + (int32_t)off;
}

}  // namespace magpo
