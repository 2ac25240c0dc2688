// Row kernels of the general Sable path (generic.cuh): the layers of rowops.cu on D-wide rows, D = 32 * VPL with VPL in {1, 2, 4}, and the
// per-(env, head) retention scan. One warp owns one row; lane l holds columns [l VPL, (l+1) VPL). Reference: networks/sable_network.py:40-343,
// networks/retention.py:66-115,229-323, networks/torsos.py:79-99; flax 0.10.3 RMSNorm / GroupNorm arithmetic;
// backward per Appendix G. These serve the non-default network shapes and favour clarity over the last GB/s: the default shape never runs them.
#include "generic.cuh"

namespace magpo {
namespace {

constexpr float kEps = 1e-6f;
constexpr int kWarps = 8;

template <int VPL>
__device__ __forceinline__ void ldv(const float* __restrict__ p, int64_t row, int ld, int lane, float (&v)[VPL]) {
  const float* q = p + row * ld + lane * VPL;
  if constexpr (VPL == 4) { const float4 t = *reinterpret_cast<const float4*>(q); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (VPL == 2) { const float2 t = *reinterpret_cast<const float2*>(q); v[0] = t.x; v[1] = t.y; }
  else v[0] = q[0];
}
template <int VPL>
__device__ __forceinline__ void stv(float* __restrict__ p, int64_t row, int ld, int lane, const float (&v)[VPL]) {
  float* q = p + row * ld + lane * VPL;
  if constexpr (VPL == 4) *reinterpret_cast<float4*>(q) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (VPL == 2) *reinterpret_cast<float2*>(q) = make_float2(v[0], v[1]);
  else q[0] = v[0];
}
__device__ __forceinline__ float swishf(float x) { return x * sigmoid_precise(x); }
__device__ __forceinline__ float swish_grad(float x) {
  const float sg = sigmoid_precise(x);
  return sg * (1.0f + x * (1.0f - sg));
}
// sum over the `span` lanes (power of two) of this lane's aligned group
__device__ __forceinline__ float seg_sum(float v, int span) {
  for (int o = span >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int VPL>
__device__ __forceinline__ float row_sumsq(const float (&p)[VPL]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s = fmaf(p[i], p[i], s);
  return warp_sum(s);
}
// per-lane column partials of all warps of the CTA -> one atomic per column; `dst_of(col)` maps a column to its slot
template <int VPL, typename F>
__device__ __forceinline__ void flush_cols(const float (&part)[VPL], float* __restrict__ dst, float* sm /*[kWarps * 32 * VPL]*/, F dst_of) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, D = 32 * VPL;
#pragma unroll
  for (int i = 0; i < VPL; ++i) sm[w * D + lane * VPL + i] = part[i];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += sm[i * D + c];
    atomicAdd(dst + dst_of(c), s);
  }
  __syncthreads();
}

inline unsigned row_grid(int64_t R) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, kWarps), (int64_t)kNumSMs * 8)); }

#define ROW_LOOP()                                                               \
  const int lane = threadIdx.x & 31;                                             \
  const int64_t wg = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);          \
  const int64_t wstride = (int64_t)gridDim.x * kWarps;                           \
  for (int64_t row = wg; row < R; row += wstride)

// ------------------------------------------------------------------ (gelu) + residual + RMSNorm (+PE)
template <int VPL>
__global__ void __launch_bounds__(256)
act_rms_fwd_k(int64_t R, const float* __restrict__ z, const float* __restrict__ res, const float* __restrict__ scale, int flags,
              const float* __restrict__ pe, const int32_t* __restrict__ step, int max_step, float* __restrict__ y, float* __restrict__ ype) {
  constexpr int D = 32 * VPL;
  float sc[VPL];
  ldv<VPL>(scale, 0, 0, threadIdx.x & 31, sc);
  ROW_LOOP() {
    float p[VPL];
    ldv<VPL>(z, row, D, lane, p);
    if (flags & ROW_GELU) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) p[i] = gelu_tanh(p[i]);
    }
    if (res) {
      float r[VPL];
      ldv<VPL>(res, row, D, lane, r);
#pragma unroll
      for (int i = 0; i < VPL; ++i) p[i] += r[i];
    }
    const float rstd = rsqrtf(row_sumsq<VPL>(p) * (1.0f / D) + kEps);
    float o[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) o[i] = p[i] * (rstd * sc[i]);
    if (y) stv<VPL>(y, row, D, lane, o);
    if (ype) {
      float e[VPL];
      ldv<VPL>(pe, min(max(step[row], 0), max_step), D, lane, e);
#pragma unroll
      for (int i = 0; i < VPL; ++i) e[i] += o[i];
      stv<VPL>(ype, row, D, lane, e);
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
act_rms_bwd_k(int64_t R, const float* __restrict__ z, const float* __restrict__ res, const float* __restrict__ scale, int flags,
              const float* __restrict__ dy1, const float* __restrict__ dy2, const float* __restrict__ dy3, float* __restrict__ dout,
              float* __restrict__ dscale) {
  constexpr int D = 32 * VPL;
  __shared__ float sm[kWarps * D];
  float sc[VPL], ds[VPL];
  ldv<VPL>(scale, 0, 0, threadIdx.x & 31, sc);
#pragma unroll
  for (int i = 0; i < VPL; ++i) ds[i] = 0.f;
  ROW_LOOP() {
    float zz[VPL], p[VPL], d[VPL];
    ldv<VPL>(z, row, D, lane, zz);
#pragma unroll
    for (int i = 0; i < VPL; ++i) p[i] = (flags & ROW_GELU) ? gelu_tanh(zz[i]) : zz[i];
    if (res) {
      float r[VPL];
      ldv<VPL>(res, row, D, lane, r);
#pragma unroll
      for (int i = 0; i < VPL; ++i) p[i] += r[i];
    }
    const float rstd = rsqrtf(row_sumsq<VPL>(p) * (1.0f / D) + kEps);
    ldv<VPL>(dy1, row, D, lane, d);
    if (dy2) { float t[VPL]; ldv<VPL>(dy2, row, D, lane, t);
#pragma unroll
      for (int i = 0; i < VPL; ++i) d[i] += t[i]; }
    if (dy3) { float t[VPL]; ldv<VPL>(dy3, row, D, lane, t);
#pragma unroll
      for (int i = 0; i < VPL; ++i) d[i] += t[i]; }
    float dotl = 0.f, u[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      ds[i] += d[i] * p[i] * rstd;
      u[i] = d[i] * sc[i];
      dotl = fmaf(p[i], u[i], dotl);
    }
    const float dot = warp_sum(dotl) * (1.0f / D);
    const float r3 = rstd * rstd * rstd;
    float dp[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      dp[i] = rstd * u[i] - p[i] * r3 * dot;
      if (flags & ROW_GELU) dp[i] *= gelu_tanh_grad(zz[i]);
    }
    stv<VPL>(dout, row, D, lane, dp);
  }
  flush_cols<VPL>(ds, dscale, sm, [](int c) { return c; });
}

// ------------------------------------------------------------------ GroupNorm (n_head groups per head row) * swish gate
// A D-wide row is n_head head rows of hs features; flax GroupNorm(num_groups = n_head) on the [rows * n_head, hs] view splits each head row
// into n_head groups of gs = hs / n_head features (retention.py:289-291). Lanes per group: gs / VPL = 32 / n_head^2.
template <int VPL>
__global__ void __launch_bounds__(256)
gn_gate_fwd_k(int64_t R, int nh, const float* __restrict__ g, int ldg, const float* __restrict__ ret, const float* __restrict__ gs,
              const float* __restrict__ gb, float* __restrict__ gated) {
  constexpr int D = 32 * VPL;
  const int hs = D / nh, span = 32 / (nh * nh);
  const float inv = 1.0f / (float)(hs / nh);
  float sc[VPL], bi[VPL];
  const int c0 = ((threadIdx.x & 31) * VPL) % hs;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { sc[i] = gs[c0 + i]; bi[i] = gb[c0 + i]; }
  ROW_LOOP() {
    float x[VPL], gg[VPL], o[VPL];
    ldv<VPL>(ret, row, D, lane, x);
    ldv<VPL>(g, row, ldg, lane, gg);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { s1 += x[i]; s2 = fmaf(x[i], x[i], s2); }
    const float mean = seg_sum(s1, span) * inv;
    const float var = fmaxf(0.0f, seg_sum(s2, span) * inv - mean * mean);  // flax "fast variance"
    const float rstd = rsqrtf(var + kEps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) o[i] = swishf(gg[i]) * ((x[i] - mean) * rstd * sc[i] + bi[i]);
    stv<VPL>(gated, row, D, lane, o);
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
gn_gate_bwd_k(int64_t R, int nh, const float* __restrict__ g, int ldg, const float* __restrict__ ret, const float* __restrict__ gs,
              const float* __restrict__ gb, const float* __restrict__ dgated, float* __restrict__ dg, int lddg, float* __restrict__ dret,
              float* __restrict__ dgs, float* __restrict__ dgb) {
  constexpr int D = 32 * VPL;
  __shared__ float sm[kWarps * D];
  const int hs = D / nh, span = 32 / (nh * nh);
  const float inv = 1.0f / (float)(hs / nh);
  float sc[VPL], bi[VPL], dS[VPL], dB[VPL];
  const int c0 = ((threadIdx.x & 31) * VPL) % hs;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { sc[i] = gs[c0 + i]; bi[i] = gb[c0 + i]; dS[i] = 0.f; dB[i] = 0.f; }
  ROW_LOOP() {
    float x[VPL], gg[VPL], dgt[VPL], xh[VPL], u[VPL], o[VPL];
    ldv<VPL>(ret, row, D, lane, x);
    ldv<VPL>(g, row, ldg, lane, gg);
    ldv<VPL>(dgated, row, D, lane, dgt);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { s1 += x[i]; s2 = fmaf(x[i], x[i], s2); }
    const float mean = seg_sum(s1, span) * inv;
    const float var = fmaxf(0.0f, seg_sum(s2, span) * inv - mean * mean);
    const float rstd = rsqrtf(var + kEps);
    float su = 0.f, sux = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xh[i] = (x[i] - mean) * rstd;
      const float nrm = xh[i] * sc[i] + bi[i];
      o[i] = dgt[i] * nrm * swish_grad(gg[i]);
      const float dn = dgt[i] * swishf(gg[i]);
      dS[i] += dn * xh[i];
      dB[i] += dn;
      u[i] = dn * sc[i];
      su += u[i];
      sux = fmaf(u[i], xh[i], sux);
    }
    stv<VPL>(dg, row, lddg, lane, o);
    const float mu = seg_sum(su, span) * inv, mux = seg_sum(sux, span) * inv;
#pragma unroll
    for (int i = 0; i < VPL; ++i) o[i] = rstd * (u[i] - mu - xh[i] * mux);
    stv<VPL>(dret, row, D, lane, o);
  }
  flush_cols<VPL>(dS, dgs, sm, [hs](int c) { return c % hs; });
  flush_cols<VPL>(dB, dgb, sm, [hs](int c) { return c % hs; });
}

// ------------------------------------------------------------------ SwiGLU middle: h = swish(gl[:, :D]) * gl[:, D:]
template <int VPL>
__global__ void __launch_bounds__(256)
swiglu_fwd_k(int64_t R, const float* __restrict__ gl, float* __restrict__ h) {
  constexpr int D = 32 * VPL;
  ROW_LOOP() {
    float a[VPL], b[VPL];
    ldv<VPL>(gl, row, 2 * D, lane, a);
    ldv<VPL>(gl + D, row, 2 * D, lane, b);
#pragma unroll
    for (int i = 0; i < VPL; ++i) a[i] = swishf(a[i]) * b[i];
    stv<VPL>(h, row, D, lane, a);
  }
}
template <int VPL>
__global__ void __launch_bounds__(256)
swiglu_bwd_k(int64_t R, const float* __restrict__ gl, const float* __restrict__ dh, float* __restrict__ dgl) {
  constexpr int D = 32 * VPL;
  ROW_LOOP() {
    float a[VPL], b[VPL], d[VPL], o1[VPL], o2[VPL];
    ldv<VPL>(gl, row, 2 * D, lane, a);
    ldv<VPL>(gl + D, row, 2 * D, lane, b);
    ldv<VPL>(dh, row, D, lane, d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) { o1[i] = d[i] * b[i] * swish_grad(a[i]); o2[i] = d[i] * swishf(a[i]); }
    stv<VPL>(dgl, row, 2 * D, lane, o1);
    stv<VPL>(dgl + D, row, 2 * D, lane, o2);
  }
}

// ------------------------------------------------------------------ decoder action embedding (decode.py:86-108)
__device__ __forceinline__ int shifted_token(const int32_t* __restrict__ action, int64_t row, int A) {
  if (A < 0) return 0;                  // inference, first agent: start-of-timestep token
  if (A == 0) return 1 + action[row];  // inference, later agents: action[] holds the previous agent's action
  return (row % A) == 0 ? 0 : 1 + action[row - 1];
}
template <int VPL>
__global__ void __launch_bounds__(256)
embed_fwd_k(int64_t R, int A, const int32_t* __restrict__ action, const float* __restrict__ Wa, const float* __restrict__ scale,
            const float* __restrict__ pe, const int32_t* __restrict__ step, int max_step, float* __restrict__ x, float* __restrict__ xpe) {
  constexpr int D = 32 * VPL;
  float sc[VPL];
  ldv<VPL>(scale, 0, 0, threadIdx.x & 31, sc);
  ROW_LOOP() {
    float p[VPL], o[VPL];
    ldv<VPL>(Wa, shifted_token(action, row, A), D, lane, p);
#pragma unroll
    for (int i = 0; i < VPL; ++i) p[i] = gelu_tanh(p[i]);
    const float rstd = rsqrtf(row_sumsq<VPL>(p) * (1.0f / D) + kEps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) o[i] = p[i] * (rstd * sc[i]);
    stv<VPL>(x, row, D, lane, o);
    if (xpe) {
      float e[VPL];
      ldv<VPL>(pe, min(max(step[row], 0), max_step), D, lane, e);
#pragma unroll
      for (int i = 0; i < VPL; ++i) e[i] += o[i];
      stv<VPL>(xpe, row, D, lane, e);
    }
  }
}
// per row: dWa[token] += J^T dy (the RMSNorm / gelu Jacobian of that token's row), dscale += dy * normalised
template <int VPL>
__global__ void __launch_bounds__(256)
embed_bwd_k(int64_t R, int A, const int32_t* __restrict__ action, const float* __restrict__ Wa, const float* __restrict__ scale,
            const float* __restrict__ dy1, const float* __restrict__ dy2, float* __restrict__ dWa, float* __restrict__ dscale) {
  constexpr int D = 32 * VPL;
  __shared__ float sm[kWarps * D];
  float sc[VPL], ds[VPL];
  ldv<VPL>(scale, 0, 0, threadIdx.x & 31, sc);
#pragma unroll
  for (int i = 0; i < VPL; ++i) ds[i] = 0.f;
  ROW_LOOP() {
    const int tok = shifted_token(action, row, A);
    float zz[VPL], p[VPL], d[VPL], u[VPL];
    ldv<VPL>(Wa, tok, D, lane, zz);
#pragma unroll
    for (int i = 0; i < VPL; ++i) p[i] = gelu_tanh(zz[i]);
    const float rstd = rsqrtf(row_sumsq<VPL>(p) * (1.0f / D) + kEps);
    ldv<VPL>(dy1, row, D, lane, d);
    if (dy2) { float t[VPL]; ldv<VPL>(dy2, row, D, lane, t);
#pragma unroll
      for (int i = 0; i < VPL; ++i) d[i] += t[i]; }
    float dotl = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { ds[i] += d[i] * p[i] * rstd; u[i] = d[i] * sc[i]; dotl = fmaf(p[i], u[i], dotl); }
    const float dot = warp_sum(dotl) * (1.0f / D);
    const float r3 = rstd * rstd * rstd;
#pragma unroll
    for (int i = 0; i < VPL; ++i) atomicAdd(dWa + (size_t)tok * D + lane * VPL + i, (rstd * u[i] - p[i] * r3 * dot) * gelu_tanh_grad(zz[i]));
  }
  flush_cols<VPL>(ds, dscale, sm, [](int c) { return c; });
}

// pe[p, 2i] = sin(p div_i), pe[p, 2i+1] = cos(p div_i), div_i = exp(2i * (-ln(10000) / D))  (positional_encoding.py:32-58)
__global__ void pe_table_k(int D, int max_step, float* __restrict__ pe) {
  const int p = blockIdx.x;
  for (int i = threadIdx.x; i < D / 2; i += blockDim.x) {
    const float div = expf((float)(2 * i) * (-logf(10000.0f) / (float)D));
    const float x = (float)p * div;
    pe[(size_t)p * D + 2 * i] = sinf(x);
    pe[(size_t)p * D + 2 * i + 1] = cosf(x);
  }
}

__global__ void __launch_bounds__(256) add_k(int64_t n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a[i] + b[i];
}

// ------------------------------------------------------------------ retention scan, one CTA per (env, head)
// H <- lam_t H (lam_t = 0 if done_t else kappa_h); encoder: H += sum_i k_i^T v_i, then o_i = q_i H for every token of the step;
// decoder: per token H += k_i^T v_i, o_i = q_i H. State [hs, hs] in shared memory (retention.py:102-115 is exactly this scan; the chunkwise
// training form :66-100 is algebraically the same, oracle test_recurrent_equals_chunkwise).
template <bool CAUSAL>
__global__ void __launch_bounds__(256)
retention_fwd_k(int T, int N, int A, int nh, int nb, int blk, int hs, const HeadKappas kappas, const float* __restrict__ q,
                const float* __restrict__ k, const float* __restrict__ v, int ld, const float* __restrict__ H0, const uint8_t* __restrict__ done,
                float* __restrict__ ret, int ldr, float* __restrict__ Hsave, float* __restrict__ Hout) {
  extern __shared__ float sm[];
  float* H = sm;                  // [hs][hs]
  float* qs = H + hs * hs;        // [A][hs]
  float* ks = qs + A * hs;
  float* vs = ks + A * hs;
  const int n = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, hh = hs * hs;
  const float kappa = kappas.k[h];
  const int64_t soff = (((int64_t)n * nh + h) * nb + blk) * hh;
  for (int i = tid; i < hh; i += blockDim.x) H[i] = H0 ? H0[soff + i] : 0.f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float lam = (done && done[(int64_t)t * N + n]) ? 0.0f : kappa;
    const int64_t base = ((int64_t)t * N + n) * A;
    for (int i = tid; i < A * hs; i += blockDim.x) {
      const int64_t off = (base + i / hs) * ld + h * hs + i % hs;
      qs[i] = q[off]; ks[i] = k[off]; vs[i] = v[off];
    }
    for (int i = tid; i < hh; i += blockDim.x) H[i] *= lam;
    __syncthreads();
    if (!CAUSAL) {
      for (int i = tid; i < hh; i += blockDim.x) {
        const int r = i / hs, c = i % hs;
        float acc = H[i];
        for (int j = 0; j < A; ++j) acc = fmaf(ks[j * hs + r], vs[j * hs + c], acc);
        H[i] = acc;
      }
      __syncthreads();
      for (int i = tid; i < A * hs; i += blockDim.x) {
        const int j = i / hs, c = i % hs;
        float acc = 0.f;
        for (int r = 0; r < hs; ++r) acc = fmaf(qs[j * hs + r], H[r * hs + c], acc);
        ret[(base + j) * ldr + h * hs + c] = acc;
      }
    } else {
      for (int j = 0; j < A; ++j) {
        for (int i = tid; i < hh; i += blockDim.x) H[i] = fmaf(ks[j * hs + i / hs], vs[j * hs + i % hs], H[i]);
        __syncthreads();
        for (int c = tid; c < hs; c += blockDim.x) {
          float acc = 0.f;
          for (int r = 0; r < hs; ++r) acc = fmaf(qs[j * hs + r], H[r * hs + c], acc);
          ret[(base + j) * ldr + h * hs + c] = acc;
        }
        __syncthreads();
      }
    }
    if (Hsave) {
      float* dst = Hsave + (((int64_t)t * N + n) * nh + h) * hh;
      for (int i = tid; i < hh; i += blockDim.x) dst[i] = H[i];
    }
    __syncthreads();
  }
  if (Hout)
    for (int i = tid; i < hh; i += blockDim.x) Hout[soff + i] = H[i];
}

// Reverse scan with G = dL/dH (state after the step, before the next step's decay). Hsave[t] = state after step t; inside a causal step
// the per-token states are recovered by exact rank-1 down-dates (as the 64-wide kernel does).
template <bool CAUSAL>
__global__ void __launch_bounds__(256)
retention_bwd_k(int T, int N, int A, int nh, int hs, const HeadKappas kappas, const float* __restrict__ q, const float* __restrict__ k,
                const float* __restrict__ v, int ld, const uint8_t* __restrict__ done, const float* __restrict__ Hsave,
                const float* __restrict__ dret, int ldr, float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int ldd) {
  extern __shared__ float sm[];
  float* H = sm;
  float* G = H + hs * hs;
  float* qs = G + hs * hs;
  float* ks = qs + A * hs;
  float* vs = ks + A * hs;
  float* ds = vs + A * hs;
  const int n = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, hh = hs * hs;
  const float kappa = kappas.k[h];
  for (int i = tid; i < hh; i += blockDim.x) G[i] = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const int64_t base = ((int64_t)t * N + n) * A;
    const float* hsrc = Hsave + (((int64_t)t * N + n) * nh + h) * hh;
    __syncthreads();
    for (int i = tid; i < hh; i += blockDim.x) H[i] = hsrc[i];
    for (int i = tid; i < A * hs; i += blockDim.x) {
      const int64_t row = base + i / hs;
      const int c = h * hs + i % hs;
      qs[i] = q[row * ld + c]; ks[i] = k[row * ld + c]; vs[i] = v[row * ld + c]; ds[i] = dret[row * ldr + c];
    }
    __syncthreads();
    if (!CAUSAL) {
      // o_j = q_j H for every token: dq_j = do_j H^T; G += sum_j q_j^T do_j; then dk_j = G v_j, dv_j = k_j^T G
      for (int i = tid; i < A * hs; i += blockDim.x) {
        const int j = i / hs, r = i % hs;
        float acc = 0.f;
        for (int c = 0; c < hs; ++c) acc = fmaf(ds[j * hs + c], H[r * hs + c], acc);
        dq[(base + j) * ldd + h * hs + r] = acc;
      }
      for (int i = tid; i < hh; i += blockDim.x) {
        const int r = i / hs, c = i % hs;
        float acc = G[i];
        for (int j = 0; j < A; ++j) acc = fmaf(qs[j * hs + r], ds[j * hs + c], acc);
        G[i] = acc;
      }
      __syncthreads();
      for (int i = tid; i < A * hs; i += blockDim.x) {
        const int j = i / hs, x = i % hs;
        float a1 = 0.f, a2 = 0.f;
        for (int c = 0; c < hs; ++c) {
          a1 = fmaf(G[x * hs + c], vs[j * hs + c], a1);   // dk_j[x]
          a2 = fmaf(ks[j * hs + c], G[c * hs + x], a2);   // dv_j[x]
        }
        dk[(base + j) * ldd + h * hs + x] = a1;
        dv[(base + j) * ldd + h * hs + x] = a2;
      }
    } else {
      for (int j = A - 1; j >= 0; --j) {
        for (int r = tid; r < hs; r += blockDim.x) {
          float acc = 0.f;
          for (int c = 0; c < hs; ++c) acc = fmaf(ds[j * hs + c], H[r * hs + c], acc);
          dq[(base + j) * ldd + h * hs + r] = acc;
        }
        for (int i = tid; i < hh; i += blockDim.x) G[i] = fmaf(qs[j * hs + i / hs], ds[j * hs + i % hs], G[i]);
        __syncthreads();
        for (int x = tid; x < hs; x += blockDim.x) {
          float a1 = 0.f, a2 = 0.f;
          for (int c = 0; c < hs; ++c) {
            a1 = fmaf(G[x * hs + c], vs[j * hs + c], a1);
            a2 = fmaf(ks[j * hs + c], G[c * hs + x], a2);
          }
          dk[(base + j) * ldd + h * hs + x] = a1;
          dv[(base + j) * ldd + h * hs + x] = a2;
        }
        for (int i = tid; i < hh; i += blockDim.x) H[i] = fmaf(-ks[j * hs + i / hs], vs[j * hs + i % hs], H[i]);  // state before token j
        __syncthreads();
      }
    }
    __syncthreads();
    const float lam = (done && done[(int64_t)t * N + n]) ? 0.0f : kappa;
    for (int i = tid; i < hh; i += blockDim.x) G[i] *= lam;
  }
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) MAGPO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MAGPO_OK;
}

}  // namespace

#define G_DISPATCH(D, CALL)                                      \
  do {                                                           \
    if ((D) == 32) { constexpr int VPL = 1; CALL; }              \
    else if ((D) == 64) { constexpr int VPL = 2; CALL; }         \
    else if ((D) == 128) { constexpr int VPL = 4; CALL; }        \
    else return MAGPO_ERR_UNSUPPORTED;                           \
    MAGPO_LAUNCH_OK();                                           \
    return MAGPO_OK;                                             \
  } while (0)

int g_act_rms_fwd(cudaStream_t s, int D, int64_t R, const float* z, const float* res, const float* scale, int flags, const float* pe,
                  const int32_t* step, int max_step, float* y, float* ype) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * D * R * (1 + (res ? 1 : 0) + (y ? 1 : 0) + (ype ? 1 : 0)));
  G_DISPATCH(D, (act_rms_fwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, z, res, scale, flags, pe, step, max_step, y, ype)));
}
int g_act_rms_bwd(cudaStream_t s, int D, int64_t R, const float* z, const float* res, const float* scale, int flags, const float* dy1,
                  const float* dy2, const float* dy3, float* dout, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * D * R * (3 + (res ? 1 : 0) + (dy2 ? 1 : 0) + (dy3 ? 1 : 0)));
  G_DISPATCH(D, (act_rms_bwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, z, res, scale, flags, dy1, dy2, dy3, dout, dscale)));
}
int g_gn_gate_fwd(cudaStream_t s, int D, int nh, int64_t R, const float* g, int ldg, const float* ret, const float* gs, const float* gb,
                  float* gated) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 12.0 * D * R);
  G_DISPATCH(D, (gn_gate_fwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, nh, g, ldg, ret, gs, gb, gated)));
}
int g_gn_gate_bwd(cudaStream_t s, int D, int nh, int64_t R, const float* g, int ldg, const float* ret, const float* gs, const float* gb,
                  const float* dgated, float* dg, int lddg, float* dret, float* dgs, float* dgb) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 20.0 * D * R);
  G_DISPATCH(D, (gn_gate_bwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, nh, g, ldg, ret, gs, gb, dgated, dg, lddg, dret, dgs, dgb)));
}
int g_swiglu_fwd(cudaStream_t s, int D, int64_t R, const float* gl, float* h) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 12.0 * D * R);
  G_DISPATCH(D, (swiglu_fwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, gl, h)));
}
int g_swiglu_bwd(cudaStream_t s, int D, int64_t R, const float* gl, const float* dh, float* dgl) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 20.0 * D * R);
  G_DISPATCH(D, (swiglu_bwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, gl, dh, dgl)));
}
int g_embed_fwd(cudaStream_t s, int D, int64_t R, int A, const int32_t* action, const float* Wa, const float* scale, const float* pe,
                const int32_t* step, int max_step, float* x, float* xpe) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 8.0 * D * R);
  G_DISPATCH(D, (embed_fwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, A, action, Wa, scale, pe, step, max_step, x, xpe)));
}
int g_embed_bwd(cudaStream_t s, int D, int64_t R, int A, const int32_t* action, const float* Wa, const float* scale, const float* dy1,
                const float* dy2, float* dWa, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 8.0 * D * R);
  G_DISPATCH(D, (embed_bwd_k<VPL><<<row_grid(R), 256, 0, s>>>(R, A, action, Wa, scale, dy1, dy2, dWa, dscale)));
}
int g_pe_table(cudaStream_t s, int D, int max_step, float* pe, bool enabled) {
  if (!enabled) {
    MAGPO_CUDA_OK(cudaMemsetAsync(pe, 0, sizeof(float) * (size_t)(max_step + 1) * D, s));
    return MAGPO_OK;
  }
  pe_table_k<<<max_step + 1, 64, 0, s>>>(D, max_step, pe);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int g_add_rows(cudaStream_t s, int64_t n, const float* a, const float* b, float* y) {
  if (n <= 0) return MAGPO_OK;
  add_k<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8), 256, 0, s>>>(n, a, b, y);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int g_retention_fwd(cudaStream_t s, const NetShape& sh, const HeadKappas& kappas, int blk, int T, int N, int rows_per_step, bool causal,
                    const float* q, const float* k, const float* v, int ld, const float* H0, const uint8_t* done, float* ret, int ldr,
                    float* Hsave, float* Hout) {
  if (T <= 0 || N <= 0) return MAGPO_OK;
  const int A = rows_per_step;
  const size_t smem = (size_t)(sh.hs * sh.hs + 3 * A * sh.hs) * sizeof(float);
  ProfScope ps(PROF_RET_FWD, s, 4.0 * 4.0 * sh.D * (double)T * N * A);
  const dim3 grid((unsigned)N, (unsigned)sh.nh);
  if (causal) {
    MAGPO_TRY(set_smem(retention_fwd_k<true>, smem));
    retention_fwd_k<true><<<grid, 256, smem, s>>>(T, N, A, sh.nh, sh.nb, blk, sh.hs, kappas, q, k, v, ld, H0, done, ret, ldr, Hsave, Hout);
  } else {
    MAGPO_TRY(set_smem(retention_fwd_k<false>, smem));
    retention_fwd_k<false><<<grid, 256, smem, s>>>(T, N, A, sh.nh, sh.nb, blk, sh.hs, kappas, q, k, v, ld, H0, done, ret, ldr, Hsave, Hout);
  }
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int g_retention_bwd(cudaStream_t s, const NetShape& sh, const HeadKappas& kappas, int T, int N, int rows_per_step, bool causal, const float* q,
                    const float* k, const float* v, int ld, const uint8_t* done, const float* Hsave, const float* dret, int ldr, float* dq,
                    float* dk, float* dv, int ldd) {
  if (T <= 0 || N <= 0) return MAGPO_OK;
  const int A = rows_per_step;
  const size_t smem = (size_t)(2 * sh.hs * sh.hs + 4 * A * sh.hs) * sizeof(float);
  ProfScope ps(PROF_RET_BWD, s, 7.0 * 4.0 * sh.D * (double)T * N * A);
  const dim3 grid((unsigned)N, (unsigned)sh.nh);
  if (causal) {
    MAGPO_TRY(set_smem(retention_bwd_k<true>, smem));
    retention_bwd_k<true><<<grid, 256, smem, s>>>(T, N, A, sh.nh, sh.hs, kappas, q, k, v, ld, done, Hsave, dret, ldr, dq, dk, dv, ldd);
  } else {
    MAGPO_TRY(set_smem(retention_bwd_k<false>, smem));
    retention_bwd_k<false><<<grid, 256, smem, s>>>(T, N, A, sh.nh, sh.hs, kappas, q, k, v, ld, done, Hsave, dret, ldr, dq, dk, dv, ldd);
  }
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo
