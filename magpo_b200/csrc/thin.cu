// Thin dense layers of the update: layers whose input width (obs_dim d <= 16) or output width (action_dim a <= 16) is a
// handful of columns.  As GEMMs they are pure streaming: one side of the product is a 16..64-byte row, so the work per
// token is a few FMAs and the kernels below are written for HBM bandwidth (one coalesced pass over the wide operand,
// weights in registers / shared memory, parameter gradients accumulated in registers and flushed once per CTA), with the
// neighbouring element-wise work fused in:
//   thin_k_fwd      Y[R,N]   = act(X[R,K] W + b)                       K <= 16, N in {64,128}   (learner pre-torso, torsos.py:36-47)
//   thin_k_bwd      dW += X^T dY, db += colsum(dY), dY optionally masked by the layer's relu output   (X is data: no dX)
//   thin_n_fwd      Y[R,N]   = X[R,128] W + b                          N <= 16                  (DiscreteActionHead, heads.py:32-63)
//   thin_n_bwd      dX = (dY W^T) * [X > 0], dW += X^T dY, db += colsum(dY)   (head backward fused with the relu of the post torso)
//   obs_embed_fwd   on = RMSNorm_d(obs); z0 = on Wobs; xin = RMSNorm(gelu(z0)); kqv = xin + PE    (sable_network.py:93-101,121-137)
//   obs_embed_bwd   dWobs += on^T dz0, dobs_scale += ...                (obs is data: no input gradient)
#include "common.cuh"
#include "kernels.cuh"

namespace magpo {
namespace {

constexpr float kEps = 1e-6f;
constexpr int kThinK = 16;  // widest "thin" side

inline unsigned thin_grid(int64_t units, int per_block) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(units, per_block), (int64_t)kNumSMs * 8));
}

// ---------------------------------------------------------------------------------------------- thin K
// Every kernel below keeps UN rows per thread/warp in flight (all loads of a batch are issued before the arithmetic): with
// 256-byte..512-byte rows the memory pipe needs ~100 rows per SM outstanding to cover the HBM latency.

// the K-wide input row of `row`, zero-padded to KMAX (one 16-byte load when the row is exactly a float4)
template <int KMAX>
__device__ __forceinline__ void load_thin_row(float (&x)[KMAX], const float* __restrict__ X, int64_t row, int ldx, int K, bool ok) {
  if (KMAX == 4 && K == 4 && ldx == 4) {
    const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(X) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[KMAX - 1] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) x[k] = (ok && k < K) ? __ldg(X + row * ldx + k) : 0.f;
  }
}

// thread = (row slot, 4 output columns); NQ = N / 4 threads per row
template <int NQ, int KMAX>
__global__ void __launch_bounds__(256)
thin_k_fwd_kernel(int64_t R, int K, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, float* __restrict__ Y, int ldy, int relu) {
  constexpr int ROWS = 256 / NQ;
  constexpr int UN = KMAX <= 4 ? 4 : 2;  // rows in flight per thread (KMAX = 16: 2 x 16 inputs next to the 16 float4 weights)
  const int c = threadIdx.x % NQ, rs = threadIdx.x / NQ;
  float4 w[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    w[k] = k < K ? *reinterpret_cast<const float4*>(W + (size_t)k * ldw + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = bias ? *reinterpret_cast<const float4*>(bias + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * ROWS;
  for (int64_t row0 = (int64_t)blockIdx.x * ROWS + rs; row0 < R; row0 += UN * stride) {
    float x[UN][KMAX];
#pragma unroll
    for (int u = 0; u < UN; ++u) load_thin_row<KMAX>(x[u], X, row0 + u * stride, ldx, K, row0 + u * stride < R);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      if (row >= R) break;
      float4 acc = b;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        acc.x = fmaf(x[u][k], w[k].x, acc.x); acc.y = fmaf(x[u][k], w[k].y, acc.y);
        acc.z = fmaf(x[u][k], w[k].z, acc.z); acc.w = fmaf(x[u][k], w[k].w, acc.w);
      }
      if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      *reinterpret_cast<float4*>(Y + row * ldy + 4 * c) = acc;
    }
  }
}

template <int NQ, int KMAX>
__global__ void __launch_bounds__(256)
thin_k_bwd_kernel(int64_t R, int K, const float* __restrict__ X, int ldx, const float* __restrict__ dY, int lddy,
                  const float* __restrict__ relu_out, float* __restrict__ dW, int lddw, float* __restrict__ db) {
  constexpr int ROWS = 256 / NQ;
  constexpr int N = 4 * NQ;
  constexpr int UN = KMAX <= 4 ? 4 : (KMAX <= 8 ? 2 : 1);
  __shared__ float red[(KMAX + 1) * N];
  const int c = threadIdx.x % NQ, rs = threadIdx.x / NQ;
  float4 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * ROWS;
  for (int64_t row0 = (int64_t)blockIdx.x * ROWS + rs; row0 < R; row0 += UN * stride) {
    float x[UN][KMAX];
    float4 d[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      const bool ok = row < R;
      d[u] = ok ? *reinterpret_cast<const float4*>(dY + row * lddy + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (relu_out && ok) {  // dY is the gradient w.r.t. relu(X W + b): mask it with the layer's saved output
        const float4 o = *reinterpret_cast<const float4*>(relu_out + row * lddy + 4 * c);
        if (!(o.x > 0.f)) d[u].x = 0.f;
        if (!(o.y > 0.f)) d[u].y = 0.f;
        if (!(o.z > 0.f)) d[u].z = 0.f;
        if (!(o.w > 0.f)) d[u].w = 0.f;
      }
      load_thin_row<KMAX>(x[u], X, row, ldx, K, ok);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      bs.x += d[u].x; bs.y += d[u].y; bs.z += d[u].z; bs.w += d[u].w;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        acc[k].x = fmaf(x[u][k], d[u].x, acc[k].x); acc[k].y = fmaf(x[u][k], d[u].y, acc[k].y);
        acc[k].z = fmaf(x[u][k], d[u].z, acc[k].z); acc[k].w = fmaf(x[u][k], d[u].w, acc[k].w);
      }
    }
  }
  for (int i = threadIdx.x; i < (KMAX + 1) * N; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
      atomicAdd(&red[k * N + 4 * c + 0], acc[k].x); atomicAdd(&red[k * N + 4 * c + 1], acc[k].y);
      atomicAdd(&red[k * N + 4 * c + 2], acc[k].z); atomicAdd(&red[k * N + 4 * c + 3], acc[k].w);
    }
  atomicAdd(&red[KMAX * N + 4 * c + 0], bs.x); atomicAdd(&red[KMAX * N + 4 * c + 1], bs.y);
  atomicAdd(&red[KMAX * N + 4 * c + 2], bs.z); atomicAdd(&red[KMAX * N + 4 * c + 3], bs.w);
  __syncthreads();
  for (int i = threadIdx.x; i < K * N; i += 256) atomicAdd(dW + (size_t)(i / N) * lddw + i % N, red[i]);
  if (db)
    for (int i = threadIdx.x; i < N; i += 256) atomicAdd(db + i, red[KMAX * N + i]);
}


// ---- N = 128 (the learner's pre-torso): a warp takes 32 consecutive rows per trip. The rows' K inputs are staged in shared memory with
// coalesced loads (the next trip's are already in registers while this trip is computed), lane l owns output columns 4l..4l+3, so every
// output row leaves / every dY row arrives as one coalesced 512-byte access and 32 rows per warp are in flight instead of one or two.
template <int KMAX>
__device__ __forceinline__ void k128_fetch(float (&v)[KMAX], const float* __restrict__ X, int ldx, int K, int64_t row0, int64_t R, int lane) {
  if (ldx == K) {  // the 32 x K block is contiguous
    const int64_t base = row0 * K, end = (row0 + 32 < R ? row0 + 32 : R) * K;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) v[i] = (i < K && base + lane + 32 * i < end) ? __ldg(X + base + lane + 32 * i) : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < KMAX; ++i) v[i] = (i < K && row0 + lane < R) ? __ldg(X + (row0 + lane) * ldx + i) : 0.f;
  }
}
template <int KMAX>
__device__ __forceinline__ void k128_stage(float* xs /*[32][KMAX]*/, const float (&v)[KMAX], int ldx, int K, int lane) {
  if (ldx == K) {
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i < K) {
        const int idx = lane + 32 * i;
        xs[(idx / K) * KMAX + idx % K] = v[i];
      }
  } else {
#pragma unroll
    for (int i = 0; i < KMAX; ++i) xs[lane * KMAX + i] = v[i];
  }
}

template <int KMAX>
__global__ void __launch_bounds__(256)
thin_k128_fwd_kernel(int64_t R, int K, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                     const float* __restrict__ bias, float* __restrict__ Y, int ldy, int relu) {
  __shared__ __align__(16) float xs_all[8][32 * KMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xs = xs_all[warp];
  float4 w[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    w[k] = k < K ? *reinterpret_cast<const float4*>(W + (size_t)k * ldw + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = bias ? *reinterpret_cast<const float4*>(bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < 32 * KMAX; i += 32) xs[i] = 0.f;  // the padding columns K..KMAX-1 stay zero
  const int64_t nblk = (R + 31) / 32, stride = (int64_t)gridDim.x * 8;
  int64_t blk = (int64_t)blockIdx.x * 8 + warp;
  float v[KMAX];
  if (blk < nblk) k128_fetch<KMAX>(v, X, ldx, K, blk * 32, R, lane);
  for (; blk < nblk; blk += stride) {
    __syncwarp();
    k128_stage<KMAX>(xs, v, ldx, K, lane);
    __syncwarp();
    if (blk + stride < nblk) k128_fetch<KMAX>(v, X, ldx, K, (blk + stride) * 32, R, lane);
    const int64_t row0 = blk * 32;
    const int nrows = (int)(R - row0 < 32 ? R - row0 : 32);
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      float4 acc = b;
#pragma unroll
      for (int q = 0; q < KMAX / 4; ++q) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + r * KMAX + 4 * q);
        const float xk[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc.x = fmaf(xk[e], w[4 * q + e].x, acc.x); acc.y = fmaf(xk[e], w[4 * q + e].y, acc.y);
          acc.z = fmaf(xk[e], w[4 * q + e].z, acc.z); acc.w = fmaf(xk[e], w[4 * q + e].w, acc.w);
        }
      }
      if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      *reinterpret_cast<float4*>(Y + (row0 + r) * ldy + 4 * lane) = acc;
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(256, 2)
thin_k128_bwd_kernel(int64_t R, int K, const float* __restrict__ X, int ldx, const float* __restrict__ dY, int lddy,
                     const float* __restrict__ relu_out, float* __restrict__ dW, int lddw, float* __restrict__ db) {
  constexpr int N = kH;
  __shared__ __align__(16) float xs_all[8][32 * KMAX];
  __shared__ float red[(KMAX + 1) * N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xs = xs_all[warp];
  float4 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < 32 * KMAX; i += 32) xs[i] = 0.f;
  const int64_t nblk = (R + 31) / 32, stride = (int64_t)gridDim.x * 8;
  for (int64_t blk = (int64_t)blockIdx.x * 8 + warp; blk < nblk; blk += stride) {
    {  // (no register prefetch here: the KMAX float4 accumulators need the registers for two CTAs per SM)
      float v[KMAX];
      k128_fetch<KMAX>(v, X, ldx, K, blk * 32, R, lane);
      __syncwarp();
      k128_stage<KMAX>(xs, v, ldx, K, lane);
      __syncwarp();
    }
    const int64_t row0 = blk * 32;
    const int nrows = (int)(R - row0 < 32 ? R - row0 : 32);
    for (int r0 = 0; r0 < nrows; r0 += 4) {
      // four rows of dY (and of the relu mask) in flight per lane before any of them is used
      float4 dq[4], oq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = r0 + u < nrows;
        dq[u] = ok ? __ldg(reinterpret_cast<const float4*>(dY + (row0 + r0 + u) * lddy + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
        oq[u] = (ok && relu_out) ? __ldg(reinterpret_cast<const float4*>(relu_out + (row0 + r0 + u) * lddy + 4 * lane))
                                 : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 d = dq[u];
        if (!(oq[u].x > 0.f)) d.x = 0.f;
        if (!(oq[u].y > 0.f)) d.y = 0.f;
        if (!(oq[u].z > 0.f)) d.z = 0.f;
        if (!(oq[u].w > 0.f)) d.w = 0.f;
        bs.x += d.x; bs.y += d.y; bs.z += d.z; bs.w += d.w;
        const int r = r0 + u < nrows ? r0 + u : 0;  // d is zero past the last row
#pragma unroll
        for (int q = 0; q < KMAX / 4; ++q) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + r * KMAX + 4 * q);
          const float xk[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[4 * q + e].x = fmaf(xk[e], d.x, acc[4 * q + e].x); acc[4 * q + e].y = fmaf(xk[e], d.y, acc[4 * q + e].y);
            acc[4 * q + e].z = fmaf(xk[e], d.z, acc[4 * q + e].z); acc[4 * q + e].w = fmaf(xk[e], d.w, acc[4 * q + e].w);
          }
        }
      }
    }
  }
  for (int i = threadIdx.x; i < (KMAX + 1) * N; i += 256) red[i] = 0.f;
  __syncthreads();
  const int c = lane;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
      atomicAdd(&red[k * N + 4 * c + 0], acc[k].x); atomicAdd(&red[k * N + 4 * c + 1], acc[k].y);
      atomicAdd(&red[k * N + 4 * c + 2], acc[k].z); atomicAdd(&red[k * N + 4 * c + 3], acc[k].w);
    }
  atomicAdd(&red[KMAX * N + 4 * c + 0], bs.x); atomicAdd(&red[KMAX * N + 4 * c + 1], bs.y);
  atomicAdd(&red[KMAX * N + 4 * c + 2], bs.z); atomicAdd(&red[KMAX * N + 4 * c + 3], bs.w);
  __syncthreads();
  for (int i = threadIdx.x; i < K * N; i += 256) atomicAdd(dW + (size_t)(i / N) * lddw + i % N, red[i]);
  if (db)
    for (int i = threadIdx.x; i < N; i += 256) atomicAdd(db + i, red[KMAX * N + i]);
}

// ---------------------------------------------------------------------------------------------- thin N (K = 128)
// warp = row; lane holds columns 4l..4l+3 of the 128-wide input. Wt (shared) is W^T padded to [NV][128].
template <int NV>
__global__ void __launch_bounds__(256)
thin_n_fwd_kernel(int64_t R, int N, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, float* __restrict__ Y, int ldy) {
  constexpr int UN = 4;
  __shared__ __align__(16) float Wt[NV * kH];
  for (int i = threadIdx.x; i < NV * kH; i += 256) {
    const int n = i / kH, k = i % kH;
    Wt[i] = n < N ? W[(size_t)k * ldw + n] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int mine = lane >> (5 - Log2<NV>::v);
  const float b = (bias && mine < N) ? bias[mine] : 0.f;
  const bool writer = (lane & ((32 >> Log2<NV>::v) - 1)) == 0 && mine < N;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), stride = (int64_t)gridDim.x * 8;
  for (int64_t row0 = w0; row0 < R; row0 += UN * stride) {
    float4 x[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      x[u] = row < R ? *reinterpret_cast<const float4*>(X + row * ldx + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      if (row >= R) break;
      float v[NV];
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float4 w = *reinterpret_cast<const float4*>(&Wt[n * kH + 4 * lane]);
        v[n] = x[u].x * w.x + x[u].y * w.y + x[u].z * w.z + x[u].w * w.w;
      }
      warp_reduce_scatter<NV>(v, lane);
      if (writer) Y[row * ldy + mine] = v[0] + b;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
thin_n_bwd_kernel(int64_t R, int N, const float* __restrict__ X, int ldx, const float* __restrict__ dY, int lddy,
                  const float* __restrict__ W, int ldw, int relu_mask, float* __restrict__ dX, int lddx,
                  float* __restrict__ dW, int lddw, float* __restrict__ db, float* __restrict__ dbx) {
  constexpr int UN = 4;
  __shared__ __align__(16) float Wt[NV * kH];
  __shared__ float red[NV * kH + NV];
  for (int i = threadIdx.x; i < NV * kH; i += 256) {
    const int n = i / kH, k = i % kH;
    Wt[i] = n < N ? W[(size_t)k * ldw + n] : 0.f;
  }
  for (int i = threadIdx.x; i < NV * kH + NV; i += 256) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float4 acc[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bs = 0.f;  // lane n: column sum of dY[:, n]
  float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of the masked dX
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), stride = (int64_t)gridDim.x * 8;
  for (int64_t row0 = w0; row0 < R; row0 += UN * stride) {
    float4 x[UN];
    float dl[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      const bool ok = row < R;
      x[u] = ok ? *reinterpret_cast<const float4*>(X + row * ldx + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      dl[u] = (ok && lane < N) ? dY[row * lddy + lane] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      bs += dl[u];
      float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float d = __shfl_sync(0xffffffffu, dl[u], n);
        const float4 w = *reinterpret_cast<const float4*>(&Wt[n * kH + 4 * lane]);
        dx.x = fmaf(d, w.x, dx.x); dx.y = fmaf(d, w.y, dx.y); dx.z = fmaf(d, w.z, dx.z); dx.w = fmaf(d, w.w, dx.w);
        acc[n].x = fmaf(d, x[u].x, acc[n].x); acc[n].y = fmaf(d, x[u].y, acc[n].y);
        acc[n].z = fmaf(d, x[u].z, acc[n].z); acc[n].w = fmaf(d, x[u].w, acc[n].w);
      }
      if (relu_mask) {
        if (!(x[u].x > 0.f)) dx.x = 0.f;
        if (!(x[u].y > 0.f)) dx.y = 0.f;
        if (!(x[u].z > 0.f)) dx.z = 0.f;
        if (!(x[u].w > 0.f)) dx.w = 0.f;
      }
      if (row < R) {
        xs.x += dx.x; xs.y += dx.y; xs.z += dx.z; xs.w += dx.w;
        if (dX) *reinterpret_cast<float4*>(dX + row * lddx + 4 * lane) = dx;
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NV; ++n)
    if (n < N) {
      atomicAdd(&red[n * kH + 4 * lane + 0], acc[n].x); atomicAdd(&red[n * kH + 4 * lane + 1], acc[n].y);
      atomicAdd(&red[n * kH + 4 * lane + 2], acc[n].z); atomicAdd(&red[n * kH + 4 * lane + 3], acc[n].w);
    }
  if (lane < N) atomicAdd(&red[NV * kH + lane], bs);
  __syncthreads();
  for (int i = threadIdx.x; i < N * kH; i += 256) atomicAdd(dW + (size_t)(i % kH) * lddw + i / kH, red[i]);
  if (db && threadIdx.x < N) atomicAdd(db + threadIdx.x, red[NV * kH + threadIdx.x]);
  if (dbx) {  // second use of the reduction buffer (first kH floats)
    __syncthreads();
    if (threadIdx.x < kH) red[threadIdx.x] = 0.f;
    __syncthreads();
    atomicAdd(&red[4 * lane + 0], xs.x); atomicAdd(&red[4 * lane + 1], xs.y);
    atomicAdd(&red[4 * lane + 2], xs.z); atomicAdd(&red[4 * lane + 3], xs.w);
    __syncthreads();
    if (threadIdx.x < kH) atomicAdd(dbx + threadIdx.x, red[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------------------------- Sable obs embedding
__device__ __forceinline__ float2 ld2(const float* __restrict__ p, int64_t row, int ld, int lane) {
  return *reinterpret_cast<const float2*>(p + row * ld + 2 * lane);
}
__device__ __forceinline__ void st2(float* __restrict__ p, int64_t row, int ld, int lane, float2 v) {
  *reinterpret_cast<float2*>(p + row * ld + 2 * lane) = v;
}

// warp = row; lane holds output columns 2l, 2l+1; the d-wide input row is read by every lane (one broadcast sector)
template <int KMAX>
__global__ void __launch_bounds__(256)
obs_embed_fwd_kernel(int64_t R, int d, const float* __restrict__ obs, const float* __restrict__ obs_scale,
                     const float* __restrict__ Wobs, const float* __restrict__ ln_scale, const float* __restrict__ pe,
                     const int32_t* __restrict__ step, int max_step, float* __restrict__ on, float* __restrict__ z0,
                     float* __restrict__ xin, float* __restrict__ kqv) {
  constexpr int UN = KMAX <= 4 ? 4 : (KMAX <= 8 ? 2 : 1);
  const int lane = threadIdx.x & 31;
  float2 w[KMAX];
  float sc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    w[k] = k < d ? *reinterpret_cast<const float2*>(Wobs + (size_t)k * kD + 2 * lane) : make_float2(0.f, 0.f);
    sc[k] = k < d ? obs_scale[k] : 0.f;
  }
  const float2 ls = *reinterpret_cast<const float2*>(ln_scale + 2 * lane);
  const float inv_d = 1.0f / (float)d;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), stride = (int64_t)gridDim.x * 8;
  // software-pipelined over rows: the next group's observation rows / step counts are in flight while this group is written
  float x[UN][KMAX], xn_[UN][KMAX];
  int st[UN], stn[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    const int64_t row = w0 + u * stride;
    load_thin_row<KMAX>(x[u], obs, row, d, d, row < R);
    st[u] = row < R ? step[row] : 0;
  }
  for (int64_t row0 = w0; row0 < R; row0 += UN * stride) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + (UN + u) * stride;
      load_thin_row<KMAX>(xn_[u], obs, row, d, d, row < R);
      stn[u] = row < R ? step[row] : 0;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = row0 + u * stride;
      if (row >= R) break;
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) ss = fmaf(x[u][k], x[u][k], ss);
      const float rstd0 = rsqrtf(ss * inv_d + kEps);
      float2 z = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const float o = x[u][k] * (rstd0 * sc[k]);
        if (on && lane == k && k < d) on[row * d + k] = o;
        z.x = fmaf(o, w[k].x, z.x);
        z.y = fmaf(o, w[k].y, z.y);
      }
      st2(z0, row, kD, lane, z);
      const float2 g = make_float2(gelu_tanh(z.x), gelu_tanh(z.y));
      const float rstd = rsqrtf(warp_sum(g.x * g.x + g.y * g.y) * (1.0f / kD) + kEps);
      const float2 o = make_float2(g.x * (rstd * ls.x), g.y * (rstd * ls.y));
      st2(xin, row, kD, lane, o);
      const float2 e = ld2(pe, min(max(st[u], 0), max_step), kD, lane);
      st2(kqv, row, kD, lane, make_float2(o.x + e.x, o.y + e.y));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      st[u] = stn[u];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) x[u][k] = xn_[u][k];
    }
  }
}

// obs_embed_bwd with the row-batch mapping of thin_k128: a warp takes 32 consecutive rows per trip, stages their observations in shared
// memory (coalesced), lane l normalises row l once (the old kernel recomputed the row's RMS in every lane and fetched the K inputs with K
// broadcast loads per row), then every row costs one coalesced 256-byte dz load and K/4 broadcast LDS.128.
template <int KMAX>
__global__ void __launch_bounds__(256, 2)
obs_embed_rows_bwd_kernel(int64_t R, int d, const float* __restrict__ obs, const float* __restrict__ obs_scale,
                          const float* __restrict__ Wobs, const float* __restrict__ dz0, float* __restrict__ dWobs,
                          float* __restrict__ dscale) {
  __shared__ __align__(16) float xs_all[8][32 * KMAX];
  __shared__ float red[KMAX * kD + KMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xs = xs_all[warp];
  float2 w[KMAX], acc[KMAX];
  float sc[KMAX], ds[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    w[k] = k < d ? *reinterpret_cast<const float2*>(Wobs + (size_t)k * kD + 2 * lane) : make_float2(0.f, 0.f);
    sc[k] = k < d ? obs_scale[k] : 0.f;
    acc[k] = make_float2(0.f, 0.f);
    ds[k] = 0.f;
  }
  for (int i = lane; i < 32 * KMAX; i += 32) xs[i] = 0.f;
  const float inv_d = 1.0f / (float)d;
  const int64_t nblk = (R + 31) / 32, stride = (int64_t)gridDim.x * 8;
  for (int64_t blk = (int64_t)blockIdx.x * 8 + warp; blk < nblk; blk += stride) {
    {
      float v[KMAX];
      k128_fetch<KMAX>(v, obs, d, d, blk * 32, R, lane);
      __syncwarp();
      k128_stage<KMAX>(xs, v, d, d, lane);
      __syncwarp();
      // lane l: row l -> x * rstd (rows past R are zero: rstd finite, products zero)
      float x[KMAX], ss = 0.f;
#pragma unroll
      for (int q = 0; q < KMAX / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(xs + lane * KMAX + 4 * q);
        x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) ss = fmaf(x[k], x[k], ss);
      const float rstd0 = rsqrtf(ss * inv_d + kEps);
#pragma unroll
      for (int q = 0; q < KMAX / 4; ++q)
        *reinterpret_cast<float4*>(xs + lane * KMAX + 4 * q) =
            make_float4(x[4 * q] * rstd0, x[4 * q + 1] * rstd0, x[4 * q + 2] * rstd0, x[4 * q + 3] * rstd0);
      __syncwarp();
    }
    const int64_t row0 = blk * 32;
    const int nrows = (int)(R - row0 < 32 ? R - row0 : 32);
    for (int r0 = 0; r0 < nrows; r0 += 4) {
      float2 dq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        dq[u] = r0 + u < nrows ? __ldg(reinterpret_cast<const float2*>(dz0 + (row0 + r0 + u) * kD + 2 * lane)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u < nrows ? r0 + u : 0;  // dz is zero past the last row
        const float2 dz = dq[u];
#pragma unroll
        for (int q = 0; q < KMAX / 4; ++q) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + r * KMAX + 4 * q);
          const float xn[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * q + e;
            const float o = xn[e] * sc[k];
            acc[k].x = fmaf(o, dz.x, acc[k].x);
            acc[k].y = fmaf(o, dz.y, acc[k].y);
            ds[k] = fmaf(dz.x * w[k].x + dz.y * w[k].y, xn[e], ds[k]);  // lane partial of d(on_k) * x_k * rstd
          }
        }
      }
    }
  }
  for (int i = threadIdx.x; i < KMAX * kD + KMAX; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < d) {
      atomicAdd(&red[k * kD + 2 * lane], acc[k].x);
      atomicAdd(&red[k * kD + 2 * lane + 1], acc[k].y);
      const float t = warp_sum(ds[k]);
      if (lane == 0) atomicAdd(&red[KMAX * kD + k], t);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < d * kD; i += 256) atomicAdd(dWobs + i, red[i]);
  if (threadIdx.x < d) atomicAdd(dscale + threadIdx.x, red[KMAX * kD + threadIdx.x]);
}

}  // namespace

bool thin_k_ok(int K, int N, int ldw, const float* W, const float* Y, int ldy) {
  return K >= 1 && K <= kThinK && (N == kD || N == kH) && !(ldw & 3) && !(ldy & 3) && !(reinterpret_cast<uintptr_t>(W) & 15) &&
         !(reinterpret_cast<uintptr_t>(Y) & 15);
}
bool thin_n_ok(int K, int N, const float* X, int ldx) {
  return N >= 1 && N <= 16 && K == kH && !(ldx & 3) && !(reinterpret_cast<uintptr_t>(X) & 15);
}

int thin_k_fwd(cudaStream_t s, int64_t R, int K, int N, const float* X, int ldx, const float* W, int ldw, const float* bias,
               float* Y, int ldy, int relu) {
  if (R <= 0) return MAGPO_OK;
  if (!thin_k_ok(K, N, ldw, W, Y, ldy) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15))) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * R * (K + N));
#define THIN_K_FWD(NQ, KM) thin_k_fwd_kernel<NQ, KM><<<thin_grid(R, 256 / NQ), 256, 0, s>>>(R, K, X, ldx, W, ldw, bias, Y, ldy, relu)
  if (N == kH) {
    const unsigned g128 = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, 256), (int64_t)kNumSMs * 3));
#define THIN_K128_FWD(KM) thin_k128_fwd_kernel<KM><<<g128, 256, 0, s>>>(R, K, X, ldx, W, ldw, bias, Y, ldy, relu)
    if (K <= 4) THIN_K128_FWD(4); else if (K <= 8) THIN_K128_FWD(8); else THIN_K128_FWD(16);
#undef THIN_K128_FWD
  }
  else { if (K <= 4) THIN_K_FWD(16, 4); else if (K <= 8) THIN_K_FWD(16, 8); else THIN_K_FWD(16, 16); }
#undef THIN_K_FWD
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int thin_k_bwd(cudaStream_t s, int64_t R, int K, int N, const float* X, int ldx, const float* dY, int lddy, const float* relu_out,
               float* dW, int lddw, float* db) {
  if (R <= 0) return MAGPO_OK;
  if (!thin_k_ok(K, N, 4, dY, dY, lddy)) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * R * (K + N));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, 64), (int64_t)kNumSMs * 4));
#define THIN_K_BWD(NQ, KM) thin_k_bwd_kernel<NQ, KM><<<grid, 256, 0, s>>>(R, K, X, ldx, dY, lddy, relu_out, dW, lddw, db)
  if (N == kH) {
    const unsigned g128 = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, 256), (int64_t)kNumSMs * 4));
#define THIN_K128_BWD(KM) thin_k128_bwd_kernel<KM><<<g128, 256, 0, s>>>(R, K, X, ldx, dY, lddy, relu_out, dW, lddw, db)
    if (K <= 4) THIN_K128_BWD(4); else if (K <= 8) THIN_K128_BWD(8); else THIN_K128_BWD(16);
#undef THIN_K128_BWD
  }
  else { if (K <= 4) THIN_K_BWD(16, 4); else if (K <= 8) THIN_K_BWD(16, 8); else THIN_K_BWD(16, 16); }
#undef THIN_K_BWD
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int thin_n_fwd(cudaStream_t s, int64_t R, int K, int N, const float* X, int ldx, const float* W, int ldw, const float* bias,
               float* Y, int ldy) {
  if (R <= 0) return MAGPO_OK;
  if (!thin_n_ok(K, N, X, ldx)) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * R * (K + N));
  if (N <= 8) thin_n_fwd_kernel<8><<<thin_grid(R, 8), 256, 0, s>>>(R, N, X, ldx, W, ldw, bias, Y, ldy);
  else thin_n_fwd_kernel<16><<<thin_grid(R, 8), 256, 0, s>>>(R, N, X, ldx, W, ldw, bias, Y, ldy);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int thin_n_bwd(cudaStream_t s, int64_t R, int K, int N, const float* X, int ldx, const float* dY, int lddy, const float* W, int ldw,
               int relu_mask, float* dX, int lddx, float* dW, int lddw, float* db, float* dbx) {
  if (R <= 0) return MAGPO_OK;
  if (!thin_n_ok(K, N, X, ldx) || (dX && ((lddx & 3) || (reinterpret_cast<uintptr_t>(dX) & 15)))) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, 4.0 * R * (2 * K + N));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, 64), (int64_t)kNumSMs * 4));
  if (N <= 8) thin_n_bwd_kernel<8><<<grid, 256, 0, s>>>(R, N, X, ldx, dY, lddy, W, ldw, relu_mask, dX, lddx, dW, lddw, db, dbx);
  else thin_n_bwd_kernel<16><<<grid, 256, 0, s>>>(R, N, X, ldx, dY, lddy, W, ldw, relu_mask, dX, lddx, dW, lddw, db, dbx);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

bool obs_embed_ok(int d) { return d >= 1 && d <= kThinK; }

int obs_embed_fwd(cudaStream_t s, int64_t R, int d, const float* obs, const float* obs_scale, const float* Wobs,
                  const float* ln_scale, const float* pe, const int32_t* step, int max_step, float* on, float* z0, float* xin,
                  float* kqv) {
  if (R <= 0) return MAGPO_OK;
  if (!obs_embed_ok(d)) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, R * (8.0 * d + 3 * 256.0));
#define OBS_FWD(KM) obs_embed_fwd_kernel<KM><<<thin_grid(R, 8), 256, 0, s>>>(R, d, obs, obs_scale, Wobs, ln_scale, pe, step, max_step, on, z0, xin, kqv)
  if (d <= 4) OBS_FWD(4); else if (d <= 8) OBS_FWD(8); else OBS_FWD(16);
#undef OBS_FWD
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int obs_embed_bwd(cudaStream_t s, int64_t R, int d, const float* obs, const float* obs_scale, const float* Wobs, const float* dz0,
                  float* dWobs, float* dscale) {
  if (R <= 0) return MAGPO_OK;
  if (!obs_embed_ok(d)) return MAGPO_ERR_UNSUPPORTED;
  ProfScope ps(PROF_ROWOPS, s, R * (4.0 * d + 256.0));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, 64), (int64_t)kNumSMs * 4));
#define OBS_BWD(KM) obs_embed_rows_bwd_kernel<KM><<<grid, 256, 0, s>>>(R, d, obs, obs_scale, Wobs, dz0, dWobs, dscale)
  if (d <= 4) OBS_BWD(4); else if (d <= 8) OBS_BWD(8); else OBS_BWD(16);
#undef OBS_BWD
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo

using namespace magpo;

// Test hook: kind 0 thin_k_fwd (relu per flags&2), 1 thin_k_bwd, 2 thin_n_fwd, 3 thin_n_bwd (relu mask per flags&2).
extern "C" int magpo_test_thin(magpo_stream_t s_, int kind, int64_t R, int K, int N, const float* X, const float* W, const float* bias,
                               const float* dY, float* out0, float* out1, float* out2, int flags) {
  cudaStream_t s = as_stream(s_);
  switch (kind) {
    case 0: return thin_k_fwd(s, R, K, N, X, K, W, N, bias, out0, N, flags & 2);
    case 1: return thin_k_bwd(s, R, K, N, X, K, dY, N, (flags & 2) ? W : nullptr, out0, N, out1);
    case 2: return thin_n_fwd(s, R, K, N, X, K, W, N, bias, out0, N);
    case 3: return thin_n_bwd(s, R, K, N, X, K, dY, N, W, N, flags & 2, out0, K, out1, N, out2, (flags & 4) ? const_cast<float*>(bias) : nullptr);
    default: return MAGPO_ERR_ARG;
  }
}
