// Reverse-scan GAE over the [T,B,A] trajectory — mava/utils/multistep.py:24-68 (calculate_gae).
// The recurrence is a dependent chain per (env, agent) column, but only its ARITHMETIC is sequential: the inputs are known up
// front. A CTA owns 32 consecutive columns (lane = column, so every load/store is a coalesced 128-byte line per warp) and its
// 4 warps own the four 32-step segments of a 128-step block of the time axis. All warps first pull their segment into registers
// (96 independent loads per thread instead of a 128-step chain of load blocks), then the segments run one after the other, latest
// first, handing the running advantage to the next through shared memory. Measured (ncu, bench size: T=128, 16 K columns, 36 MB):
// 22 us, 1.06 TB/s of DRAM traffic — at this size the kernel is one launch + one DRAM round trip + the four dependent segments,
// not bandwidth; the single-pass version it replaces took 23-33 us. Arithmetic is the un-fused fp32 sequence of
// the reference (delta = r + g*nv*(1-nd) - v; gae = delta + (g*l)*(1-nd)*gae), so results are bit-identical to the NumPy oracle.
// Algorithmic traffic: 17 B per agent-step.
#include "common.cuh"

namespace magpo {

constexpr int kGaeSeg = 32;    // timesteps per warp segment
constexpr int kGaeWarps = 4;   // segments per block of the time axis

__global__ void __launch_bounds__(kGaeWarps * 32, 4)
gae_kernel(int T, int B, int A, const float* __restrict__ reward, const float* __restrict__ value,
           const uint8_t* __restrict__ done, const float* __restrict__ last_value,
           const uint8_t* __restrict__ last_done, float gamma, float gl, float* __restrict__ adv,
           float* __restrict__ targets) {
  __shared__ float carry[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t cols = (int64_t)B * A;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const bool live = c < cols;
  const int b = live ? (int)(c / A) : 0;
  if (w == 0) carry[lane] = 0.0f;
  constexpr int kBlock = kGaeSeg * kGaeWarps;
  for (int base = ((T - 1) / kBlock) * kBlock; base >= 0; base -= kBlock) {
    const int s0 = base + w * kGaeSeg, s1 = min(s0 + kGaeSeg, T);  // this warp's timesteps [s0, s1), possibly empty
    float r[kGaeSeg], v[kGaeSeg];
    uint8_t d[kGaeSeg];
    float nv = 0.0f, nd = 0.0f;
    if (live && s0 < s1) {
#pragma unroll
      for (int i = 0; i < kGaeSeg; ++i) {
        if (s0 + i < s1) {
          const int64_t off = (int64_t)(s0 + i) * cols + c;
          r[i] = __ldg(reward + off);
          v[i] = __ldg(value + off);
          d[i] = __ldg(done + (int64_t)(s0 + i) * B + b);
        }
      }
      // value / done that follow the segment's last step
      nv = s1 < T ? __ldg(value + (int64_t)s1 * cols + c) : last_value[c];
      nd = (s1 < T ? __ldg(done + (int64_t)s1 * B + b) : last_done[b]) ? 1.0f : 0.0f;
    }
    for (int seg = kGaeWarps - 1; seg >= 0; --seg) {
      __syncthreads();  // the carry of the later segment (or of the previous block / the initial 0) is visible
      if (w == seg && live && s0 < s1) {
        float acc = carry[lane];
#pragma unroll
        for (int i = kGaeSeg - 1; i >= 0; --i) {
          if (s0 + i < s1) {
            const float nnd = __fsub_rn(1.0f, nd);
            const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(__fmul_rn(gamma, nv), nnd)), v[i]);
            acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnd), acc));
            const int64_t off = (int64_t)(s0 + i) * cols + c;
            adv[off] = acc;
            targets[off] = __fadd_rn(acc, v[i]);
            nv = v[i];
            nd = d[i] ? 1.0f : 0.0f;
          }
        }
        carry[lane] = acc;
      }
    }
    __syncthreads();
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
  gae_kernel<<<(unsigned)ceil_div(cols, 32), kGaeWarps * 32, 0, as_stream(s)>>>(
      T, B, A, reward, value, done, last_value, last_done, (float)gamma, (float)(gamma * gae_lambda), advantages, targets);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
