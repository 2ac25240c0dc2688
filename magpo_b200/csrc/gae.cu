// Reverse-scan GAE over the [T,B,A] trajectory — mava/utils/multistep.py:24-68 (calculate_gae).
// One thread per (env, agent) column walks t = T-1..0; consecutive threads own consecutive columns,
// so every load/store is a coalesced 128-byte line per warp. Loads of a block of timesteps are
// issued before the dependent FMA chain consumes them. Arithmetic is the un-fused fp32 sequence
// of the reference (delta = r + g*nv*(1-nd) - v; gae = delta + (g*l)*(1-nd)*gae), so results are
// bit-identical to the NumPy oracle. Algorithmic traffic: 17 B per agent-step.
#include "common.cuh"

namespace magpo {

constexpr int kGaeUnroll = 8;

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
  int t = T - 1;
  while (t >= 0) {
    float r[kGaeUnroll], v[kGaeUnroll];
    uint8_t d[kGaeUnroll];
    const int n = min(kGaeUnroll, t + 1);
#pragma unroll
    for (int u = 0; u < kGaeUnroll; ++u) {
      if (u < n) {
        const int64_t off = (int64_t)(t - u) * cols + c;
        r[u] = __ldg(reward + off);
        v[u] = __ldg(value + off);
        d[u] = __ldg(done + (int64_t)(t - u) * B + b);
      }
    }
#pragma unroll
    for (int u = 0; u < kGaeUnroll; ++u) {
      if (u < n) {
        const float nnd = __fsub_rn(1.0f, nd);
        const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(__fmul_rn(gamma, nv), nnd)), v[u]);
        acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnd), acc));
        const int64_t off = (int64_t)(t - u) * cols + c;
        adv[off] = acc;
        targets[off] = __fadd_rn(acc, v[u]);
        nv = v[u];
        nd = d[u] ? 1.0f : 0.0f;
      }
    }
    t -= n;
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
