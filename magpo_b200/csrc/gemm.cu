// fp32 SIMT GEMMs for the tall-skinny shapes of the rec_magpo nets (rows = tokens, K/N <= 384).
//   gemm_nn : Y[M,N] = X[M,K] @ W[K,N] (+ bias) (+ Y)          forward layers and dX = dY @ W^T (W^T pre-transposed)
//   gemm_tn : dW[K,N] += X[M,K]^T @ dY[M,N]                    weight gradients (row-slab split, fp32 atomics)
//   colsum  : db[N]  += sum_m dY[m,N]                          bias gradients
// These carry the strict-fp32 path (rollout inference, where sampled actions must not flip, and the
// parity baseline of the update); the tcgen05 3xTF32 path in gemm_tc.cu takes over the large update GEMMs.
#include "common.cuh"
#include "kernels.cuh"

namespace magpo {

namespace {
constexpr int BM = 128, BN = 64, BK = 32, TM = 8, TN = 4;
constexpr int XS_LD = BM + 4;  // k-major X tile, padded, keeps 16-byte alignment of the 8-row reads

__global__ void __launch_bounds__(256)
gemm_nn_kernel(int M, int N, int K, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
               const float* __restrict__ bias, float* __restrict__ Y, int ldy, int flags) {
  __shared__ __align__(16) float Xs[BK * XS_LD];
  __shared__ __align__(16) float Ws[BK * BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
  const bool xvec = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    // X tile: 128 rows x 32 k, stored k-major
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + i * 256;
      const int row = f >> 3, kq = (f & 7) * 4;
      const int64_t m = m0 + row;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < M) {
        const float* src = X + m * ldx + k0 + kq;
        if (xvec && k0 + kq + 3 < K) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k0 + kq + j < K) v[j] = __ldg(src + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Xs[(kq + j) * XS_LD + row] = v[j];
    }
    // W tile: 32 k x 64 n
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = tid + i * 256;
      const int k = f >> 6, n = f & 63;
      float v = 0.f;
      if (k0 + k < K && n0 + n < N) v = __ldg(W + (int64_t)(k0 + k) * ldw + n0 + n);
      Ws[k * BN + n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[k * XS_LD + ty * TM]);
      const float4 xb = *reinterpret_cast<const float4*>(&Xs[k * XS_LD + ty * TM + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k * BN + tx * TN]);
      const float xr[TM] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const float wr[TN] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool accumulate = flags & GEMM_ACCUMULATE;
  const bool relu = flags & GEMM_RELU;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      float* dst = Y + m * ldy + n;
      if (accumulate) v += *dst;
      if (relu) v = fmaxf(v, 0.0f);
      *dst = v;
    }
  }
}

// dW[kt*64.., nt*64..] += sum over this CTA's row slab of X^T dY. 256 threads, 4x4 outputs each.
constexpr int TB_M = 32;
__global__ void __launch_bounds__(256)
gemm_tn_kernel(int64_t M, int N, int K, const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy,
               float* __restrict__ dW, int ldw, int64_t rows_per_cta) {
  __shared__ __align__(16) float Xs[TB_M * 64];
  __shared__ __align__(16) float Ys[TB_M * 64];
  const int tid = threadIdx.x;
  const int tn = tid & 15, tk = tid >> 4;
  const int k0 = blockIdx.z * 64, n0 = blockIdx.y * 64;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t me = min(M, mb + rows_per_cta);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t m0 = mb; m0 < me; m0 += TB_M) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = tid + i * 256;
      const int r = f >> 6, c = f & 63;
      const int64_t m = m0 + r;
      float xv = 0.f, yv = 0.f;
      if (m < me) {
        if (k0 + c < K) xv = __ldg(X + m * ldx + k0 + c);
        if (n0 + c < N) yv = __ldg(dY + m * ldy + n0 + c);
      }
      Xs[r * 64 + c] = xv;
      Ys[r * 64 + c] = yv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < TB_M; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&Xs[r * 64 + tk * 4]);
      const float4 y = *reinterpret_cast<const float4*>(&Ys[r * 64 + tn * 4]);
      const float xr[4] = {x.x, x.y, x.z, x.w};
      const float yr[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], yr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < N) atomicAdd(dW + (int64_t)k * ldw + n, acc[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256)
colsum_kernel(int64_t M, int N, const float* __restrict__ dY, int ldy, float* __restrict__ db, int64_t rows_per_cta) {
  __shared__ float part[4][64];
  const int c = threadIdx.x & 63, rq = threadIdx.x >> 6;
  const int n = blockIdx.y * 64 + c;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_cta, me = min(M, mb + rows_per_cta);
  float s = 0.f;
  if (n < N)
    for (int64_t m = mb + rq; m < me; m += 4) s += __ldg(dY + m * ldy + n);
  part[rq][c] = s;
  __syncthreads();
  if (rq == 0 && n < N) atomicAdd(db + n, part[0][c] + part[1][c] + part[2][c] + part[3][c]);
}

__global__ void transpose_kernel(int R, int Cc, const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? in[(int64_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[(int64_t)c * R + r] = tile[threadIdx.x][i];
  }
}
}  // namespace

int gemm_nn(cudaStream_t s, int64_t M, int N, int K, const float* X, int ldx, Wref Wr, const float* bias, float* Y,
            int ldy, int flags) {
  if (M <= 0 || N <= 0) return MAGPO_OK;
  if (K <= 0) return MAGPO_ERR_ARG;
  if (Wr.wt && K >= 256 && K % 128 == 0 && !bias && !(flags & GEMM_RELU) && tc_supported(M, N, 128, X, ldx, Y, ldy, Wr.ldwt)) {
    // a [K > 256, N] weight (hi + lo images) does not fit in shared memory beside the pipeline: run K in slices of 128 that
    // accumulate into Y (TMA reduce-add) instead of falling back to 32-column tiles that re-read X four times
    for (int k0 = 0; k0 < K; k0 += 128)
      MAGPO_TRY(gemm_nn(s, M, N, 128, X + k0, ldx, wref(Wr.w + (size_t)k0 * Wr.ldw, Wr.ldw, Wr.wt + k0, Wr.ldwt), nullptr, Y, ldy,
                        k0 ? (flags | GEMM_ACCUMULATE) : flags));
    return MAGPO_OK;
  }
  if (Wr.wt && !((flags & GEMM_ACCUMULATE) && ((flags & GEMM_RELU) || bias)) && tc_supported(M, N, K, X, ldx, Y, ldy, Wr.ldwt)) {
    const float *hi, *lo;
    if (tc_lookup(Wr.wt, &hi, &lo)) return gemm_tc(s, M, N, K, X, ldx, hi, lo, Wr.ldwt, bias, Y, ldy, flags);
  }
  const float* W = Wr.w;
  const int ldw = Wr.ldw;
  ProfScope ps(M < 65536 ? PROF_GEMM_SMALL : PROF_GEMM_NN, s, 2.0 * (double)M * N * K, 4.0 * (double)M * (K + N));
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
  gemm_nn_kernel<<<grid, 256, 0, s>>>((int)M, N, K, X, ldx, W, ldw, bias, Y, ldy, flags);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

static int64_t slab_rows(int64_t M, int tiles) {
  // aim for ~4 CTAs per SM over all output tiles; slabs are multiples of 32 rows
  int64_t want = std::max<int64_t>(1, (int64_t)(4 * kNumSMs) / max(1, tiles));
  int64_t rows = ceil_div(M, want);
  rows = ceil_div(rows, TB_M) * TB_M;
  return std::max<int64_t>(rows, TB_M);
}

int gemm_tn(cudaStream_t s, int64_t M, int N, int K, const float* X, int ldx, const float* dY, int ldy, float* dW,
            int ldw) {
  if (M <= 0 || N <= 0 || K <= 0) return MAGPO_OK;
  if (tc_tn_supported(M, N, K, X, ldx, dY, ldy, dW, ldw)) return gemm_tc_tn(s, M, N, K, X, ldx, dY, ldy, dW, ldw);
  ProfScope ps(PROF_GEMM_TN, s, 2.0 * (double)M * N * K, 4.0 * (double)M * (K + N));
  const int nt = (int)ceil_div(N, 64), kt = (int)ceil_div(K, 64);
  const int64_t rows = slab_rows(M, nt * kt);
  dim3 grid((unsigned)ceil_div(M, rows), nt, kt);
  gemm_tn_kernel<<<grid, 256, 0, s>>>(M, N, K, X, ldx, dY, ldy, dW, ldw, rows);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int colsum(cudaStream_t s, int64_t M, int N, const float* dY, int ldy, float* db) {
  if (M <= 0 || N <= 0) return MAGPO_OK;
  ProfScope ps(PROF_COLSUM, s, 4.0 * (double)M * N);
  const int nt = (int)ceil_div(N, 64);
  int64_t rows = std::max<int64_t>(64, ceil_div(M, (int64_t)(4 * kNumSMs) / nt + 1));
  dim3 grid((unsigned)ceil_div(M, rows), nt);
  colsum_kernel<<<grid, 256, 0, s>>>(M, N, dY, ldy, db, rows);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int transpose(cudaStream_t s, int R, int Cc, const float* in, float* out) {
  dim3 grid((unsigned)ceil_div(Cc, 32), (unsigned)ceil_div(R, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(R, Cc, in, out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo

using namespace magpo;

// Test hooks (exported so that tests/ can check the building blocks against NumPy through the C ABI).
extern "C" int magpo_test_gemm(magpo_stream_t s, int kind, int64_t M, int N, int K, const float* A, const float* B,
                               const float* bias, float* Cout, int flags) {
  cudaStream_t st = as_stream(s);
  if (kind == 0) return gemm_nn(st, M, N, K, A, K, wref(B, N), bias, Cout, N, flags);
  if (kind == 1) return gemm_tn(st, M, N, K, A, K, B, N, Cout, N);
  if (kind == 2) return colsum(st, M, N, A, N, Cout);
  if (kind == 3) return transpose(st, (int)M, N, A, Cout);
  return MAGPO_ERR_ARG;
}
