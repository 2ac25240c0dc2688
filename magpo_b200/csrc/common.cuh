// Shared helpers for the magpo_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>

#include "../../include/magpo_b200.h"

namespace magpo {

void set_cuda_error(cudaError_t e, const char* file, int line);

#define MAGPO_CUDA_OK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::magpo::set_cuda_error(_e, __FILE__, __LINE__);        \
      return MAGPO_ERR_CUDA;                                  \
    }                                                         \
  } while (0)

void note_launch();
#define MAGPO_LAUNCH_OK()              \
  do {                                 \
    ::magpo::note_launch();            \
    MAGPO_CUDA_OK(cudaGetLastError()); \
  } while (0)

// Optional per-category device timing (CUDA events on the launching stream), used by bench.py for the live
// roofline numbers and the per-kernel breakdown. Off by default; zero cost when off.
enum ProfCat {
  PROF_GEMM_NN = 0, PROF_GEMM_TN, PROF_COLSUM, PROF_ROWOPS, PROF_RET_FWD, PROF_RET_BWD, PROF_GRU, PROF_LOSS, PROF_PACK,
  PROF_OPTIM, PROF_ENV, PROF_SAMPLE, PROF_GAE, PROF_MISC, PROF_GEMM_SMALL /* M < 64 Ki rows: the rollout */, PROF_NUM
};
struct ProfScope {
  int idx;
  cudaStream_t s;
  ProfScope(int cat, cudaStream_t s, double work, double bytes = 0.0);  // bytes: algorithmic HBM bytes when `work` is flops
  ~ProfScope();
};

#define MAGPO_TRY(expr)          \
  do {                           \
    int _r = (expr);             \
    if (_r != MAGPO_OK) return _r; \
  } while (0)

constexpr int kNumSMs = 148;
constexpr float kF32Min = -3.4028234663852886e+38f;  // jnp.finfo(float32).min

inline cudaStream_t as_stream(magpo_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over the caller-provided workspace (256-byte aligned slices).
struct Arena {
  char* base;
  size_t cap, off;
  bool overflow;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0), overflow(false) {}
  template <typename T>
  T* get(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (off + bytes > cap) {
      overflow = true;
      off += bytes;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Sum of NV per-lane values across the warp (NV a power of two): afterwards v[0] of lane l holds the warp total of value
// index l >> (5 - log2 NV). NV - 1 + (5 - log2 NV) shuffles instead of 5 NV.
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane) {
  int bit = 16;
#pragma unroll
  for (int half = NV / 2; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = lane & bit;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  for (; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
}
template <int NV> struct Log2 { static constexpr int v = 1 + Log2<NV / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
// flax nn.gelu(approximate=True)
__device__ __forceinline__ float gelu_tanh(float x) {
  const float c = 0.7978845608028654f;
  float t = tanhf(c * (x + 0.044715f * x * x * x));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float c = 0.7978845608028654f;
  float x2 = x * x;
  float t = tanhf(c * (x + 0.044715f * x * x2));
  float dt = (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * dt;
}

}  // namespace magpo
