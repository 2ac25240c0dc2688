// jax.random entry points (threefry2x32, partitionable) — SURVEY.md Appendix A1-A5.
// Replaces the jax.random call sites of rec_magpo.py:135,202,373,439,443,450,642,660,699.
#include "common.cuh"
#include "prng.cuh"

namespace magpo {

__global__ void split_kernel(const uint32_t* __restrict__ key, int num, uint32_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  uint32_t o0, o1;
  prng_split_i(key[0], key[1], (uint32_t)i, o0, o1);
  out[2 * i] = o0;
  out[2 * i + 1] = o1;
}

__global__ void bits_kernel(const uint32_t* __restrict__ key, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = prng_bits_i(key[0], key[1], (uint64_t)i);
}

__global__ void randint_kernel(const uint32_t* __restrict__ key, int64_t n, int minval, int maxval,
                               int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a0, a1, b0, b1;
  prng_split_i(key[0], key[1], 0u, a0, a1);  // k1, k2 = split(key)
  prng_split_i(key[0], key[1], 1u, b0, b1);
  uint32_t hi = prng_bits_i(a0, a1, (uint64_t)i);
  uint32_t lo = prng_bits_i(b0, b1, (uint64_t)i);
  out[i] = prng_randint_from_bits(hi, lo, minval, maxval);
}

__global__ void gumbel_kernel(const uint32_t* __restrict__ key, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = prng_gumbel_from_bits(prng_bits_i(key[0], key[1], (uint64_t)i));
}

// One round of _shuffle: sort_keys = random_bits(subkey_r, (n,)), where (key, subkey) = split(key)
// has been applied r+1 times to the input key.
__global__ void perm_keys_kernel(const uint32_t* __restrict__ key, int round, int n, uint32_t* __restrict__ sort_keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k0 = key[0], k1 = key[1], s0 = 0, s1 = 0;
  for (int r = 0; r <= round; ++r) {
    uint32_t n0, n1;
    prng_split_i(k0, k1, 1u, s0, s1);
    prng_split_i(k0, k1, 0u, n0, n1);
    k0 = n0; k1 = n1;
  }
  sort_keys[i] = prng_bits_i(s0, s1, (uint64_t)i);
}

// Stable sort by counting ranks: rank_i = #{j: key_j < key_i or (key_j == key_i and j < i)}.
__global__ void perm_rank_scatter_kernel(const uint32_t* __restrict__ sort_keys, const int32_t* __restrict__ x_in,
                                         int n, int first_round, int32_t* __restrict__ x_out) {
  __shared__ uint32_t tile[256];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t ki = i < n ? sort_keys[i] : 0u;
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    int j = base + threadIdx.x;
    tile[threadIdx.x] = j < n ? sort_keys[j] : 0xFFFFFFFFu;
    __syncthreads();
    int lim = min(256, n - base);
    for (int t = 0; t < lim; ++t) {
      uint32_t kj = tile[t];
      rank += (kj < ki) || (kj == ki && (base + t) < i);
    }
    __syncthreads();
  }
  if (i < n) x_out[rank] = first_round ? i : x_in[i];
}

int permutation_rounds(int n) {
  if (n <= 1) return 0;
  double r = 3.0 * log((double)n) / log(4294967295.0);
  return (int)ceil(r);
}

__global__ void iota_kernel(int n, int32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}


// ---------------------------------------------------------------- jax.random.normal / truncated_normal (parameter initialisers)
// jax 0.6.0 `_normal_real`: u = uniform(key, shape, f32, minval = nextafter(-1, 0), maxval = 1); sqrt(2) * erf_inv(u), and
// `_truncated_normal`: u = uniform(key, minval = erf(lower / sqrt2), maxval = erf(upper / sqrt2)); clip(sqrt(2) * erf_inv(u), ...).
// uniform: bits >> 9 | 0x3F800000 -> [1, 2) - 1 -> * (maxval - minval) + minval -> max(minval, .). erf_inv is XLA's single-precision
// expansion (two degree-8 polynomials in w = -log1p(-x^2), Giles 2010), restated here; the libm log1pf / sqrtf may differ from XLA's
// own in the last bit.
__device__ __forceinline__ float xla_erf_inv_f32(float x) {
  float w = -log1pf(-x * x);
  const bool lt = w < 5.0f;
  w = lt ? w - 2.5f : sqrtf(w) - 3.0f;
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
  p = (lt ? 3.43273939e-07f : 0.000100950558f) + p * w;
  p = (lt ? -3.5233877e-06f : 0.00134934322f) + p * w;
  p = (lt ? -4.39150654e-06f : -0.00367342844f) + p * w;
  p = (lt ? 0.00021858087f : 0.00573950773f) + p * w;
  p = (lt ? -0.00125372503f : -0.0076224613f) + p * w;
  p = (lt ? -0.00417768164f : 0.00943887047f) + p * w;
  p = (lt ? 0.246640727f : 1.00167406f) + p * w;
  p = (lt ? 1.50140941f : 2.83297682f) + p * w;
  return fabsf(x) == 1.0f ? x * INFINITY : p * x;
}
__device__ __forceinline__ float uniform_range_from_bits(uint32_t bits, float minval, float maxval) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  return fmaxf(minval, f * (maxval - minval) + minval);
}
__global__ void __launch_bounds__(256)
normal_kernel(const uint32_t* __restrict__ key, int64_t n, float minval, float maxval, float clip_lo, float clip_hi, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = uniform_range_from_bits(prng_bits_i(key[0], key[1], (uint64_t)i), minval, maxval);
  const float v = 1.41421356237309504880f * xla_erf_inv_f32(u);
  out[i] = fminf(fmaxf(v, clip_lo), clip_hi);
}

}  // namespace magpo

using namespace magpo;

extern "C" {

int magpo_prng_split(magpo_stream_t s, const uint32_t* key, int32_t num, uint32_t* out) {
  if (!key || !out || num < 0) return MAGPO_ERR_ARG;
  if (num == 0) return MAGPO_OK;
  split_kernel<<<(unsigned)ceil_div(num, 256), 256, 0, as_stream(s)>>>(key, num, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_prng_random_bits(magpo_stream_t s, const uint32_t* key, int64_t n, uint32_t* out) {
  if (!key || !out || n < 0) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  bits_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(key, n, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_prng_randint(magpo_stream_t s, const uint32_t* key, int64_t n, int32_t minval, int32_t maxval, int32_t* out) {
  if (!key || !out || n < 0) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  randint_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(key, n, minval, maxval, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_prng_gumbel(magpo_stream_t s, const uint32_t* key, int64_t n, float* out) {
  if (!key || !out || n < 0) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  gumbel_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(key, n, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_prng_permutation(magpo_stream_t s, const uint32_t* key, int32_t n, int32_t* out, uint32_t* scratch) {
  if (!key || !out || !scratch || n < 0) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  cudaStream_t st = as_stream(s);
  unsigned g = (unsigned)ceil_div(n, 256);
  int rounds = permutation_rounds(n);
  if (rounds == 0) {
    iota_kernel<<<g, 256, 0, st>>>(n, out);
    MAGPO_LAUNCH_OK();
    return MAGPO_OK;
  }
  uint32_t* sort_keys = scratch;
  int32_t* tmp = reinterpret_cast<int32_t*>(scratch + n);
  // ping-pong so that the last round lands in `out`
  int32_t* bufs[2] = {out, tmp};
  int cur = (rounds % 2 == 1) ? 0 : 1;  // destination of round 0
  const int32_t* src = nullptr;
  for (int r = 0; r < rounds; ++r) {
    perm_keys_kernel<<<g, 256, 0, st>>>(key, r, n, sort_keys);
    MAGPO_LAUNCH_OK();
    perm_rank_scatter_kernel<<<g, 256, 0, st>>>(sort_keys, src, n, r == 0, bufs[cur]);
    MAGPO_LAUNCH_OK();
    src = bufs[cur];
    cur ^= 1;
  }
  return MAGPO_OK;
}

// jax.random.normal(key, (n,), float32)
int magpo_prng_normal(magpo_stream_t s, const uint32_t* key, int64_t n, float* out) {
  if (!key || !out || n < 0) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  normal_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(key, n, nextafterf(-1.0f, 0.0f), 1.0f, -INFINITY, INFINITY, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

// jax.random.truncated_normal(key, lower, upper, (n,), float32)
int magpo_prng_truncated_normal(magpo_stream_t s, const uint32_t* key, int64_t n, float lower, float upper, float* out) {
  if (!key || !out || n < 0 || !(lower < upper)) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  const float sqrt2 = 1.41421356237309504880f;
  normal_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(key, n, erff(lower / sqrt2), erff(upper / sqrt2),
                                                                    nextafterf(lower, INFINITY), nextafterf(upper, -INFINITY), out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

// jax.random.fold_in(key, data) = threefry2x32(key, (0, data)) — a host function (keys are host values while networks are initialised)
int magpo_prng_fold_in_host(const uint32_t* key, uint32_t data, uint32_t* out) {
  if (!key || !out) return MAGPO_ERR_ARG;
  threefry2x32(key[0], key[1], 0u, data, out[0], out[1]);
  return MAGPO_OK;
}

}  // extern "C"
