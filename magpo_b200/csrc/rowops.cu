// Row-wise (per-token) layers of the Sable guider, forward and hand-derived backward:
// RMSNorm (+gelu, +residual, +positional encoding), GroupNorm*swish gate, SwiGLU middle, the heads'
// gelu->RMSNorm->Dense tail, the decoder's action embedding.  Reference: networks/sable_network.py:40-343,
// networks/retention.py:289-295, networks/torsos.py:79-99, utils/sable/positional_encoding.py:24-58,
// flax 0.10.3 RMSNorm/GroupNorm arithmetic, backward per Appendix G.
// One warp owns one 64-wide row (lane l holds columns 2l, 2l+1 -> 256-byte coalesced rows); per-column
// parameter gradients are accumulated in registers over a grid-stride loop and flushed once per CTA.
// All of these are HBM-bound: algorithmic traffic = (inputs + outputs) * 256 B per token.
#include "common.cuh"
#include "kernels.cuh"

namespace magpo {
namespace {

constexpr float kEps = 1e-6f;
constexpr int kWarps = 8;

__device__ __forceinline__ float2 ld2(const float* __restrict__ p, int64_t row, int ld, int lane) {
  return *reinterpret_cast<const float2*>(p + row * ld + 2 * lane);
}
__device__ __forceinline__ void st2(float* __restrict__ p, int64_t row, int ld, int lane, float2 v) {
  *reinterpret_cast<float2*>(p + row * ld + 2 * lane) = v;
}
__device__ __forceinline__ float swishf(float x) { return x * sigmoid_precise(x); }
// The backward kernels run in the update only (never in the rollout, whose sampled actions must not move): ex2.approx / rcp.approx forms,
// absolute error ~1e-7, a third of the instructions of expf / tanhf + IEEE division.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
// swish(x) and swish'(x) from one sigmoid
__device__ __forceinline__ void swish_both_fast(float x, float& s, float& ds) {
  const float sg = sigmoid_fast(x);
  s = x * sg;
  ds = sg * (1.0f + x * (1.0f - sg));
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float c = 0.7978845608028654f;
  const float x2 = x * x;
  const float t = tanh_fast(c * (x + 0.044715f * x * x2));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x2);
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float c = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanh_fast(c * (x + 0.044715f * x * x * x)));
}

// Flush per-lane column partials (2 columns per lane) of all warps of the CTA into global with one atomic per column.
__device__ __forceinline__ void flush_cols(float2 part, float* __restrict__ dst, float* sm /*[kWarps*64]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  sm[w * 64 + 2 * lane] = part.x;
  sm[w * 64 + 2 * lane + 1] = part.y;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += sm[i * 64 + threadIdx.x];
    atomicAdd(dst + threadIdx.x, s);
  }
  __syncthreads();
}

inline unsigned row_grid(int64_t R) { return (unsigned)std::min<int64_t>(ceil_div(R, kWarps), (int64_t)kNumSMs * 8); }

#define ROW_LOOP()                                                               \
  const int lane = threadIdx.x & 31;                                             \
  const int64_t wg = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);          \
  const int64_t wstride = (int64_t)gridDim.x * kWarps;                           \
  for (int64_t row = wg; row < R; row += wstride)

// ------------------------------------------------------------------ general-width RMSNorm (obs encoder input)
__global__ void __launch_bounds__(256)
rms_general_fwd_kernel(int64_t R, int C, const float* __restrict__ x, const float* __restrict__ scale,
                       float* __restrict__ y) {
  ROW_LOOP() {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = x[row * C + c];
      ss += v * v;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)C + kEps);
    for (int c = lane; c < C; c += 32) y[row * C + c] = x[row * C + c] * (rstd * scale[c]);
  }
}

__global__ void __launch_bounds__(256)
rms_general_bwd_scale_kernel(int64_t R, int C, const float* __restrict__ x, const float* __restrict__ dy,
                             float* __restrict__ dscale) {
  extern __shared__ float sm[];  // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int nper = (C + 31) / 32;
  float part[4] = {0.f, 0.f, 0.f, 0.f};  // C <= 128
  ROW_LOOP() {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = x[row * C + c];
      ss += v * v;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)C + kEps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (i < nper && c < C) part[i] += dy[row * C + c] * x[row * C + c] * rstd;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (threadIdx.x & 31) + 32 * i;
    if (c < C) atomicAdd(&sm[c], part[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dscale + c, sm[c]);
}

// ------------------------------------------------------------------ (gelu) + residual + RMSNorm (+PE)
__global__ void __launch_bounds__(256)
act_rms_fwd_kernel(int64_t R, const float* __restrict__ z, const float* __restrict__ res,
                   const float* __restrict__ scale, int flags, const float* __restrict__ pe,
                   const int32_t* __restrict__ step, int max_step, float* __restrict__ y, float* __restrict__ ype) {
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * (threadIdx.x & 31));
  ROW_LOOP() {
    float2 p = ld2(z, row, kD, lane);
    if (flags & ROW_GELU) { p.x = gelu_tanh(p.x); p.y = gelu_tanh(p.y); }
    if (res) { const float2 r = ld2(res, row, kD, lane); p.x += r.x; p.y += r.y; }
    const float ss = warp_sum(p.x * p.x + p.y * p.y);
    const float rstd = rsqrtf(ss * (1.0f / kD) + kEps);
    float2 o = make_float2(p.x * (rstd * sc.x), p.y * (rstd * sc.y));
    if (y) st2(y, row, kD, lane, o);
    if (ype) {
      const int st = min(max(step[row], 0), max_step);
      const float2 e = ld2(pe, st, kD, lane);
      st2(ype, row, kD, lane, make_float2(o.x + e.x, o.y + e.y));
    }
  }
}

__global__ void __launch_bounds__(256)
act_rms_bwd_kernel(int64_t R, const float* __restrict__ z, const float* __restrict__ res,
                   const float* __restrict__ scale, int flags, const float* __restrict__ dy1,
                   const float* __restrict__ dy2, const float* __restrict__ dy3, float* __restrict__ dout,
                   float* __restrict__ dscale) {
  __shared__ float sm[kWarps * 64];
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * (threadIdx.x & 31));
  float2 ds = make_float2(0.f, 0.f);
  ROW_LOOP() {
    const float2 zz = ld2(z, row, kD, lane);
    float2 p = zz;
    if (flags & ROW_GELU) { p.x = gelu_fast(p.x); p.y = gelu_fast(p.y); }
    if (res) { const float2 r = ld2(res, row, kD, lane); p.x += r.x; p.y += r.y; }
    const float ss = warp_sum(p.x * p.x + p.y * p.y);
    const float rstd = rsqrtf(ss * (1.0f / kD) + kEps);
    float2 d = ld2(dy1, row, kD, lane);
    if (dy2) { const float2 t = ld2(dy2, row, kD, lane); d.x += t.x; d.y += t.y; }
    if (dy3) { const float2 t = ld2(dy3, row, kD, lane); d.x += t.x; d.y += t.y; }
    ds.x += d.x * p.x * rstd;
    ds.y += d.y * p.y * rstd;
    const float2 u = make_float2(d.x * sc.x, d.y * sc.y);
    const float dot = warp_sum(p.x * u.x + p.y * u.y) * (1.0f / kD);
    const float r3 = rstd * rstd * rstd;
    float2 dp = make_float2(rstd * u.x - p.x * r3 * dot, rstd * u.y - p.y * r3 * dot);
    if (flags & ROW_GELU) { dp.x *= gelu_grad_fast(zz.x); dp.y *= gelu_grad_fast(zz.y); }
    st2(dout, row, kD, lane, dp);
  }
  flush_cols(ds, dscale, sm);
}

// ------------------------------------------------------------------ GroupNorm(1 group) * swish gate
__device__ __forceinline__ void ln_stats(float2 x, float& mean, float& rstd) {
  mean = warp_sum(x.x + x.y) * (1.0f / kD);
  const float m2 = warp_sum(x.x * x.x + x.y * x.y) * (1.0f / kD);
  const float var = fmaxf(0.0f, m2 - mean * mean);  // flax "fast variance"
  rstd = rsqrtf(var + kEps);
}

__global__ void __launch_bounds__(256)
gn_gate_fwd_kernel(int64_t R, const float* __restrict__ g, int ldg, const float* __restrict__ ret,
                   const float* __restrict__ gs, const float* __restrict__ gb, float* __restrict__ gated) {
  const float2 sc = *reinterpret_cast<const float2*>(gs + 2 * (threadIdx.x & 31));
  const float2 bi = *reinterpret_cast<const float2*>(gb + 2 * (threadIdx.x & 31));
  ROW_LOOP() {
    const float2 x = ld2(ret, row, kD, lane);
    const float2 gg = ld2(g, row, ldg, lane);
    float mean, rstd;
    ln_stats(x, mean, rstd);
    const float nx = (x.x - mean) * rstd * sc.x + bi.x;
    const float ny = (x.y - mean) * rstd * sc.y + bi.y;
    st2(gated, row, kD, lane, make_float2(swishf(gg.x) * nx, swishf(gg.y) * ny));
  }
}

__global__ void __launch_bounds__(256)
gn_gate_bwd_kernel(int64_t R, const float* __restrict__ g, int ldg, const float* __restrict__ ret,
                   const float* __restrict__ gs, const float* __restrict__ gb, const float* __restrict__ dgated,
                   float* __restrict__ dg, int lddg, float* __restrict__ dret, float* __restrict__ dgs,
                   float* __restrict__ dgb) {
  __shared__ float sm[kWarps * 64];
  const float2 sc = *reinterpret_cast<const float2*>(gs + 2 * (threadIdx.x & 31));
  const float2 bi = *reinterpret_cast<const float2*>(gb + 2 * (threadIdx.x & 31));
  float2 dS = make_float2(0.f, 0.f), dB = make_float2(0.f, 0.f);
  ROW_LOOP() {
    const float2 x = ld2(ret, row, kD, lane);
    const float2 gg = ld2(g, row, ldg, lane);
    const float2 dgt = ld2(dgated, row, kD, lane);
    float mean, rstd;
    ln_stats(x, mean, rstd);
    const float2 xh = make_float2((x.x - mean) * rstd, (x.y - mean) * rstd);
    const float2 nrm = make_float2(xh.x * sc.x + bi.x, xh.y * sc.y + bi.y);
    float2 sw, dsw;
    swish_both_fast(gg.x, sw.x, dsw.x);
    swish_both_fast(gg.y, sw.y, dsw.y);
    st2(dg, row, lddg, lane, make_float2(dgt.x * nrm.x * dsw.x, dgt.y * nrm.y * dsw.y));
    const float2 dn = make_float2(dgt.x * sw.x, dgt.y * sw.y);
    dS.x += dn.x * xh.x; dS.y += dn.y * xh.y;
    dB.x += dn.x; dB.y += dn.y;
    const float2 u = make_float2(dn.x * sc.x, dn.y * sc.y);
    const float mu = warp_sum(u.x + u.y) * (1.0f / kD);
    const float mux = warp_sum(u.x * xh.x + u.y * xh.y) * (1.0f / kD);
    st2(dret, row, kD, lane, make_float2(rstd * (u.x - mu - xh.x * mux), rstd * (u.y - mu - xh.y * mux)));
  }
  flush_cols(dS, dgs, sm);
  flush_cols(dB, dgb, sm);
}

// ------------------------------------------------------------------ SwiGLU middle
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(int64_t R, const float* __restrict__ gl, float* __restrict__ h) {
  ROW_LOOP() {
    const float2 a = ld2(gl, row, 2 * kD, lane);
    const float2 b = ld2(gl + kD, row, 2 * kD, lane);
    st2(h, row, kD, lane, make_float2(swishf(a.x) * b.x, swishf(a.y) * b.y));
  }
}
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(int64_t R, const float* __restrict__ gl, const float* __restrict__ dh, float* __restrict__ dgl) {
  ROW_LOOP() {
    const float2 a = ld2(gl, row, 2 * kD, lane);
    const float2 b = ld2(gl + kD, row, 2 * kD, lane);
    const float2 d = ld2(dh, row, kD, lane);
    float2 sw, dsw;
    swish_both_fast(a.x, sw.x, dsw.x);
    swish_both_fast(a.y, sw.y, dsw.y);
    st2(dgl, row, 2 * kD, lane, make_float2(d.x * b.x * dsw.x, d.y * b.y * dsw.y));
    st2(dgl + kD, row, 2 * kD, lane, make_float2(d.x * sw.x, d.y * sw.y));
  }
}

// ------------------------------------------------------------------ head tail: gelu -> RMSNorm -> Dense(nout)
// gelu(x) and gelu'(x) from one tanh
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
  const float c = 0.7978845608028654f;
  const float x2 = x * x;
  const float t = tanh_fast(c * (x + 0.044715f * x * x2));  // head_bwd only: update path
  g = 0.5f * x * (1.0f + t);
  dg = 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x2);
}


// The last Dense's weights live in registers: lane l holds rows 2l, 2l+1 of W3 [64, nout] (padded to NV columns).
template <int NV>
__global__ void __launch_bounds__(256)
head_fwd_kernel(int64_t R, const float* __restrict__ zh, const float* __restrict__ scale,
                const float* __restrict__ W3, const float* __restrict__ b3, int nout, float* __restrict__ out) {
  const int lane_ = threadIdx.x & 31;
  float w0[NV], w1[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    w0[j] = j < nout ? W3[(2 * lane_) * nout + j] : 0.f;
    w1[j] = j < nout ? W3[(2 * lane_ + 1) * nout + j] : 0.f;
  }
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * lane_);
  const int mine = lane_ >> (5 - Log2<NV>::v);
  const bool writer = (lane_ & ((32 >> Log2<NV>::v) - 1)) == 0 && mine < nout;
  const float bias = mine < nout ? b3[mine] : 0.f;
  constexpr int UN = 4;  // rows in flight per warp
  const int lane = lane_;
  const int64_t wg = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), wstride = (int64_t)gridDim.x * kWarps;
  for (int64_t row0 = wg; row0 < R; row0 += UN * wstride) {
    float2 zz[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) zz[u] = row0 + u * wstride < R ? ld2(zh, row0 + u * wstride, kD, lane) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * wstride;
      if (row >= R) break;
      const float2 p = make_float2(gelu_tanh(zz[u].x), gelu_tanh(zz[u].y));
      const float ss = warp_sum(p.x * p.x + p.y * p.y);
      const float rstd = rsqrtf(ss * (1.0f / kD) + kEps);
      const float2 hn = make_float2(p.x * (rstd * sc.x), p.y * (rstd * sc.y));
      float v[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = fmaf(hn.x, w0[j], hn.y * w1[j]);
      warp_reduce_scatter<NV>(v, lane);
      if (writer) out[row * nout + mine] = v[0] + bias;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
head_bwd_kernel(int64_t R, const float* __restrict__ zh, const float* __restrict__ scale,
                const float* __restrict__ W3, int nout, const float* __restrict__ dout, float* __restrict__ dzh,
                float* __restrict__ dscale, float* __restrict__ dW3, float* __restrict__ db3, float* __restrict__ dbz) {
  __shared__ float dWs[kD * NV];
  __shared__ float sm[kWarps * 64];
  __shared__ float dbs[NV];
  for (int i = threadIdx.x; i < kD * NV; i += blockDim.x) dWs[i] = 0.f;
  if (threadIdx.x < NV) dbs[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane_ = threadIdx.x & 31;
  float w0[NV], w1[NV], dwx[NV], dwy[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    w0[j] = j < nout ? W3[(2 * lane_) * nout + j] : 0.f;
    w1[j] = j < nout ? W3[(2 * lane_ + 1) * nout + j] : 0.f;
    dwx[j] = 0.f; dwy[j] = 0.f;
  }
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * lane_);
  float2 ds = make_float2(0.f, 0.f), dz = make_float2(0.f, 0.f);
  float dbl = 0.f;
  constexpr int UN = NV <= 8 ? 4 : 2;  // rows in flight per warp
  const int lane = lane_;
  const int64_t wg = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), wstride = (int64_t)gridDim.x * kWarps;
  for (int64_t row0 = wg; row0 < R; row0 += UN * wstride) {
    float2 zs[UN];
    float dls[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * wstride;
      const bool ok = row < R;
      zs[u] = ok ? ld2(zh, row, kD, lane) : make_float2(0.f, 0.f);
      dls[u] = (ok && lane < nout) ? dout[row * nout + lane] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * wstride;
      if (row >= R) break;
      const float2 zz = zs[u];
      float2 p, gp;
      gelu_both(zz.x, p.x, gp.x);
      gelu_both(zz.y, p.y, gp.y);
      const float ss = warp_sum(p.x * p.x + p.y * p.y);
      const float rstd = rsqrtf(ss * (1.0f / kD) + kEps);
      const float2 hn = make_float2(p.x * (rstd * sc.x), p.y * (rstd * sc.y));
      const float dol = dls[u];
      dbl += dol;
      float2 dhn = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float dj = __shfl_sync(0xffffffffu, dol, j);  // zero for j >= nout
        dhn.x = fmaf(dj, w0[j], dhn.x);
        dhn.y = fmaf(dj, w1[j], dhn.y);
        dwx[j] = fmaf(hn.x, dj, dwx[j]);
        dwy[j] = fmaf(hn.y, dj, dwy[j]);
      }
      ds.x += dhn.x * p.x * rstd;
      ds.y += dhn.y * p.y * rstd;
      const float2 u2 = make_float2(dhn.x * sc.x, dhn.y * sc.y);
      const float dot = warp_sum(p.x * u2.x + p.y * u2.y) * (1.0f / kD);
      const float r3 = rstd * rstd * rstd;
      const float2 dzv = make_float2((rstd * u2.x - p.x * r3 * dot) * gp.x, (rstd * u2.y - p.y * r3 * dot) * gp.y);
      dz.x += dzv.x; dz.y += dzv.y;  // column sums of dzh = bias gradient of the Dense that produced zh
      st2(dzh, row, kD, lane, dzv);
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (j < nout) {
      atomicAdd(&dWs[(2 * lane_) * nout + j], dwx[j]);
      atomicAdd(&dWs[(2 * lane_ + 1) * nout + j], dwy[j]);
    }
  }
  if (lane_ < nout) atomicAdd(&dbs[lane_], dbl);
  __syncthreads();
  for (int i = threadIdx.x; i < kD * nout; i += blockDim.x) atomicAdd(dW3 + i, dWs[i]);
  if (threadIdx.x < nout) atomicAdd(db3 + threadIdx.x, dbs[threadIdx.x]);
  flush_cols(ds, dscale, sm);
  if (dbz) flush_cols(dz, dbz, sm);
}

// ------------------------------------------------------------------ decoder action embedding
__device__ __forceinline__ int shifted_token(const int32_t* __restrict__ action, int64_t row, int A) {
  if (A < 0) return 0;                  // inference, first agent: start-of-timestep token
  if (A == 0) return 1 + action[row];  // inference, later agents: action[] holds the previous agent's action
  return (row % A) == 0 ? 0 : 1 + action[row - 1];
}

__global__ void __launch_bounds__(256)
embed_fwd_kernel(int64_t R, int A, const int32_t* __restrict__ action, const float* __restrict__ Wa,
                 const float* __restrict__ scale, const float* __restrict__ pe, const int32_t* __restrict__ step,
                 int max_step, float* __restrict__ x, float* __restrict__ xpe) {
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * (threadIdx.x & 31));
  ROW_LOOP() {
    const int tok = shifted_token(action, row, A);
    float2 p = ld2(Wa, tok, kD, lane);
    p.x = gelu_tanh(p.x); p.y = gelu_tanh(p.y);
    const float ss = warp_sum(p.x * p.x + p.y * p.y);
    const float rstd = rsqrtf(ss * (1.0f / kD) + kEps);
    const float2 o = make_float2(p.x * (rstd * sc.x), p.y * (rstd * sc.y));
    st2(x, row, kD, lane, o);
    const int st = min(max(step[row], 0), max_step);
    const float2 e = ld2(pe, st, kD, lane);
    st2(xpe, row, kD, lane, make_float2(o.x + e.x, o.y + e.y));
  }
}

// The map dy -> dWa[token] is linear in dy for a fixed token (the RMSNorm / gelu Jacobian depends on Wa[token] only), so
// the rows are first summed per token (a streaming pass: one or two 256-byte loads and two shared-memory adds per row,
// four rows in flight per warp) and the Jacobian is applied once per token and block.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(int64_t R, int A, int a, const int32_t* __restrict__ action, const float* __restrict__ Wa,
                 const float* __restrict__ scale, const float* __restrict__ dy1, const float* __restrict__ dy2,
                 float* __restrict__ dWa, float* __restrict__ dscale) {
  __shared__ float D[(kMaxActions + 1) * kD];
  __shared__ float sm[kWarps * 64];
  for (int i = threadIdx.x; i < (a + 1) * kD; i += blockDim.x) D[i] = 0.f;
  __syncthreads();
  constexpr int UN = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wg = (int64_t)blockIdx.x * kWarps + warp, wstride = (int64_t)gridDim.x * kWarps;
  for (int64_t row0 = wg; row0 < R; row0 += UN * wstride) {
    float2 d[UN];
    int tok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * wstride;
      const bool ok = row < R;
      tok[u] = ok ? shifted_token(action, row, A) : 0;
      d[u] = ok ? ld2(dy1, row, kD, lane) : make_float2(0.f, 0.f);
      if (ok && dy2) { const float2 t = ld2(dy2, row, kD, lane); d[u].x += t.x; d[u].y += t.y; }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (row0 + u * wstride < R) {
        atomicAdd(&D[tok[u] * kD + 2 * lane], d[u].x);
        atomicAdd(&D[tok[u] * kD + 2 * lane + 1], d[u].y);
      }
  }
  __syncthreads();
  const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * lane);
  float2 ds = make_float2(0.f, 0.f);
  for (int t = warp; t <= a; t += kWarps) {
    const float2 zz = ld2(Wa, t, kD, lane);
    const float2 p = make_float2(gelu_tanh(zz.x), gelu_tanh(zz.y));
    const float rstd = rsqrtf(warp_sum(p.x * p.x + p.y * p.y) * (1.0f / kD) + kEps);
    const float2 dd = *reinterpret_cast<const float2*>(&D[t * kD + 2 * lane]);
    ds.x += dd.x * p.x * rstd;
    ds.y += dd.y * p.y * rstd;
    const float2 u = make_float2(dd.x * sc.x, dd.y * sc.y);
    const float dot = warp_sum(p.x * u.x + p.y * u.y) * (1.0f / kD);
    const float r3 = rstd * rstd * rstd;
    atomicAdd(dWa + t * kD + 2 * lane, (rstd * u.x - p.x * r3 * dot) * gelu_tanh_grad(zz.x));
    atomicAdd(dWa + t * kD + 2 * lane + 1, (rstd * u.y - p.y * r3 * dot) * gelu_tanh_grad(zz.y));
  }
  flush_cols(ds, dscale, sm);
}

__global__ void __launch_bounds__(256)
add_pe_kernel(int64_t R, const float* __restrict__ x, const float* __restrict__ pe, const int32_t* __restrict__ step,
              int max_step, float* __restrict__ y) {
  ROW_LOOP() {
    const float2 v = ld2(x, row, kD, lane);
    const int st = min(max(step[row], 0), max_step);
    const float2 e = ld2(pe, st, kD, lane);
    st2(y, row, kD, lane, make_float2(v.x + e.x, v.y + e.y));
  }
}

// pe[p, 2i] = sin(p * div_i), pe[p, 2i+1] = cos(p * div_i), div_i = exp(2i * (-ln(10000)/64))  (positional_encoding.py:32-58)
__global__ void pe_table_kernel(int max_step, float* __restrict__ pe) {
  const int p = blockIdx.x, i = threadIdx.x;  // i in [0, 32)
  if (p > max_step) return;
  const float div = expf((float)(2 * i) * (-logf(10000.0f) / (float)kD));
  const float x = (float)p * div;
  pe[p * kD + 2 * i] = sinf(x);
  pe[p * kD + 2 * i + 1] = cosf(x);
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

int rms_general_fwd(cudaStream_t s, int64_t R, int C, const float* x, const float* scale, float* y) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 8.0 * R * C);
  rms_general_fwd_kernel<<<row_grid(R), 256, 0, s>>>(R, C, x, scale, y);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int rms_general_bwd_scale(cudaStream_t s, int64_t R, int C, const float* x, const float* dy, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 8.0 * R * C);
  if (C > 128) return MAGPO_ERR_UNSUPPORTED;
  rms_general_bwd_scale_kernel<<<row_grid(R), 256, C * sizeof(float), s>>>(R, C, x, dy, dscale);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int act_rms_fwd(cudaStream_t s, int64_t R, const float* z, const float* res, const float* scale, int flags,
                const float* pe, const int32_t* step, int max_step, float* y, float* ype) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (res ? 3 : 2) * 256.0 * R + (ype ? 256.0 * R : 0));
  act_rms_fwd_kernel<<<row_grid(R), 256, 0, s>>>(R, z, res, scale, flags, pe, step, max_step, y, ype);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int act_rms_bwd(cudaStream_t s, int64_t R, const float* z, const float* res, const float* scale, int flags,
                const float* dy1, const float* dy2, const float* dy3, float* dout, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (2 + (res ? 1 : 0) + 1 + (dy2 ? 1 : 0) + (dy3 ? 1 : 0)) * 256.0 * R);
  act_rms_bwd_kernel<<<row_grid(R), 256, 0, s>>>(R, z, res, scale, flags, dy1, dy2, dy3, dout, dscale);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int gn_gate_fwd(cudaStream_t s, int64_t R, const float* g, int ldg, const float* ret, const float* gn_scale,
                const float* gn_bias, float* gated) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 3 * 256.0 * R);
  gn_gate_fwd_kernel<<<row_grid(R), 256, 0, s>>>(R, g, ldg, ret, gn_scale, gn_bias, gated);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int gn_gate_bwd(cudaStream_t s, int64_t R, const float* g, int ldg, const float* ret, const float* gn_scale,
                const float* gn_bias, const float* dgated, float* dg, int lddg, float* dret, float* dgn_scale,
                float* dgn_bias) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 5 * 256.0 * R);
  gn_gate_bwd_kernel<<<row_grid(R), 256, 0, s>>>(R, g, ldg, ret, gn_scale, gn_bias, dgated, dg, lddg, dret,
                                                 dgn_scale, dgn_bias);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int swiglu_fwd(cudaStream_t s, int64_t R, const float* gl, float* h) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 3 * 256.0 * R);
  swiglu_fwd_kernel<<<row_grid(R), 256, 0, s>>>(R, gl, h);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int swiglu_bwd(cudaStream_t s, int64_t R, const float* gl, const float* dh, float* dgl) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 5 * 256.0 * R);
  swiglu_bwd_kernel<<<row_grid(R), 256, 0, s>>>(R, gl, dh, dgl);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int head_fwd(cudaStream_t s, int64_t R, const float* zh, const float* scale, const float* W3, const float* b3,
             int nout, float* out) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (256.0 + 4.0 * nout) * R);
  if (nout < 1 || nout > kMaxActions) return MAGPO_ERR_UNSUPPORTED;
  if (nout == 1) head_fwd_kernel<1><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, b3, nout, out);
  else if (nout <= 8) head_fwd_kernel<8><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, b3, nout, out);
  else if (nout <= 16) head_fwd_kernel<16><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, b3, nout, out);
  else head_fwd_kernel<32><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, b3, nout, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int head_bwd(cudaStream_t s, int64_t R, const float* zh, const float* scale, const float* W3, int nout,
             const float* dout, float* dzh, float* dscale, float* dW3, float* db3, float* dbz) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (512.0 + 4.0 * nout) * R);
  if (nout < 1 || nout > kMaxActions) return MAGPO_ERR_UNSUPPORTED;
  if (nout == 1) head_bwd_kernel<1><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, nout, dout, dzh, dscale, dW3, db3, dbz);
  else if (nout <= 8) head_bwd_kernel<8><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, nout, dout, dzh, dscale, dW3, db3, dbz);
  else if (nout <= 16) head_bwd_kernel<16><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, nout, dout, dzh, dscale, dW3, db3, dbz);
  else head_bwd_kernel<32><<<row_grid(R), 256, 0, s>>>(R, zh, scale, W3, nout, dout, dzh, dscale, dW3, db3, dbz);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int embed_fwd(cudaStream_t s, int64_t R, int A, const int32_t* action, const float* Wa, const float* scale,
              const float* pe, const int32_t* step, int max_step, float* x, float* xpe) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (512.0 + 8.0) * R);
  embed_fwd_kernel<<<row_grid(R), 256, 0, s>>>(R, A, action, Wa, scale, pe, step, max_step, x, xpe);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int embed_bwd(cudaStream_t s, int64_t R, int A, int a, const int32_t* action, const float* Wa, const float* scale,
              const float* dy1, const float* dy2, float* dWa, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, (256.0 + (dy2 ? 256.0 : 0) + 8.0) * R);
  if (a > kMaxActions) return MAGPO_ERR_UNSUPPORTED;
  embed_bwd_kernel<<<row_grid(R), 256, 0, s>>>(R, A, a, action, Wa, scale, dy1, dy2, dWa, dscale);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int add_pe(cudaStream_t s, int64_t R, const float* x, const float* pe, const int32_t* step, int max_step, float* y) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_ROWOPS, s, 512.0 * R);
  add_pe_kernel<<<row_grid(R), 256, 0, s>>>(R, x, pe, step, max_step, y);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int build_pe_table(cudaStream_t s, int max_step, float* pe, bool enabled) {
  if (!enabled) {
    MAGPO_CUDA_OK(cudaMemsetAsync(pe, 0, sizeof(float) * (size_t)(max_step + 1) * kD, s));
    return MAGPO_OK;
  }
  pe_table_kernel<<<max_step + 1, 32, 0, s>>>(max_step, pe);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}
int fill_f32(cudaStream_t s, float* p, int64_t n, float v) {
  if (n <= 0) return MAGPO_OK;
  fill_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 16), 256, 0, s>>>(p, n, v);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo
