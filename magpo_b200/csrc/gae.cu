// Reverse-scan GAE over the [T,B,A] trajectory — mava/utils/multistep.py:24-68 (calculate_gae).
// One thread per (env, agent) column walks t = T-1..0; consecutive threads own consecutive columns,
// so every load/store is a coalesced 128-byte line per warp. Loads run two blocks of timesteps
// ahead of the dependent chain that consumes them. Arithmetic is the un-fused fp32 sequence
// of the reference (delta = r + g*nv*(1-nd) - v; gae = delta + (g*l)*(1-nd)*gae), so results are
// bit-identical to the NumPy oracle. Algorithmic traffic: 17 B per agent-step.
#include "common.cuh"

namespace magpo {

constexpr int kGaeUnroll = 16;

struct GaeBlock {
  float r[kGaeUnroll], v[kGaeUnroll];
  uint8_t d[kGaeUnroll];
};

// loads of timesteps t, t-1, ..., t-kGaeUnroll+1 of column c (those >= 0)
__device__ __forceinline__ void gae_load(GaeBlock& k, int t, int64_t cols, int64_t c, int B, int b, const float* __restrict__ reward,
                                         const float* __restrict__ value, const uint8_t* __restrict__ done) {
#pragma unroll
  for (int u = 0; u < kGaeUnroll; ++u) {
    if (t - u >= 0) {
      const int64_t off = (int64_t)(t - u) * cols + c;
      k.r[u] = __ldg(reward + off);
      k.v[u] = __ldg(value + off);
      k.d[u] = __ldg(done + (int64_t)(t - u) * B + b);
    }
  }
}

// The columns of a step are few (B*A = 16 K at the bench size: 3.5 warps per SM), so the kernel lives on the bytes each thread keeps
// in flight: two blocks of 16 timesteps are ping-ponged, the loads of the next block are issued before the dependent chain of
// the current one runs (96 loads in flight per thread).
__global__ void __launch_bounds__(128)
gae_kernel(int T, int B, int A, const float* __restrict__ reward, const float* __restrict__ value,
           const uint8_t* __restrict__ done, const float* __restrict__ last_value,
           const uint8_t* __restrict__ last_done, float gamma, float gl, float* __restrict__ adv,
           float* __restrict__ targets) {
  const int64_t cols = (int64_t)B * A;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int b = (int)(c / A);
  float acc = 0.0f;
  float nv = last_value[c];
  float nd = last_done[b] ? 1.0f : 0.0f;
  auto consume = [&](const GaeBlock& k, int t) {
#pragma unroll
    for (int u = 0; u < kGaeUnroll; ++u) {
      if (t - u >= 0) {
        const float nnd = __fsub_rn(1.0f, nd);
        const float delta = __fsub_rn(__fadd_rn(k.r[u], __fmul_rn(__fmul_rn(gamma, nv), nnd)), k.v[u]);
        acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnd), acc));
        const int64_t off = (int64_t)(t - u) * cols + c;
        adv[off] = acc;
        targets[off] = __fadd_rn(acc, k.v[u]);
        nv = k.v[u];
        nd = k.d[u] ? 1.0f : 0.0f;
      }
    }
  };
  GaeBlock k0, k1;
  int t = T - 1;
  gae_load(k0, t, cols, c, B, b, reward, value, done);
  while (t >= 0) {
    gae_load(k1, t - kGaeUnroll, cols, c, B, b, reward, value, done);
    consume(k0, t);
    t -= kGaeUnroll;
    if (t < 0) break;
    gae_load(k0, t - kGaeUnroll, cols, c, B, b, reward, value, done);
    consume(k1, t);
    t -= kGaeUnroll;
  }
}

}  // namespace magpo

using namespace magpo;

extern "C" int magpo_gae(magpo_stream_t s, int32_t T, int32_t B, int32_t A, const float* reward, const float* value,
                         const uint8_t* done, const float* last_value, const uint8_t* last_done, double gamma,
                         double gae_lambda, float* advantages, float* targets) {
  if (T < 0 || B < 0 || A < 1 || !reward || !value || !done || !last_value || !last_done || !advantages || !targets)
    return MAGPO_ERR_ARG;
  if (T == 0 || B == 0) return MAGPO_OK;
  const int64_t cols = (int64_t)B * A;
  ProfScope ps(PROF_GAE, as_stream(s), 17.0 * (double)T * cols + 5.0 * cols);
  gae_kernel<<<(unsigned)ceil_div(cols, 128), 128, 0, as_stream(s)>>>(
      T, B, A, reward, value, done, last_value, last_done, (float)gamma, (float)(gamma * gae_lambda), advantages, targets);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
