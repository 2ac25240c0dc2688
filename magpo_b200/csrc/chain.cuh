// Building blocks of the fused row-chain kernels (chain_fwd.cu, chain_bwd.cu): a persistent CTA walks 128-row tiles of the token
// batch through a whole chain of row-local layers [row op -> GEMM -> row op -> GEMM ...] without leaving the SM.
//   * thread = row: the 128 threads of a "row group" (4 warps = the 4 TMEM sub-partitions) own one token row each, 64 floats in
//     registers; RMSNorm / GroupNorm statistics are in-thread reductions, no shuffles
//   * the A operand of every GEMM is written by its row thread straight into TENSOR MEMORY (tcgen05.st, lane = row) as the TF32 hi / lo
//     pair of the 3xTF32 product, and tcgen05.mma reads it from there (A-from-TMEM form); the accumulator comes back with tcgen05.ld.
//     Activations never pass through shared memory between two layers.
//   * the B operands (transposed weights, hi / lo images) are loaded once per CTA by TMA and stay resident in shared memory
//   * inputs arrive through a TMA ring of [128 rows x 32 floats] 128B-swizzled boxes, outputs leave through TMA stores
//   * two row groups per CTA work on alternate tiles so that one group's arithmetic hides the other's MMA / TMEM / TMA latency
#pragma once
#include "tc_ptx.cuh"

namespace magpo {
namespace chain {
using namespace tcp;

constexpr int CH_THREADS = 384;            // warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare; 4-7 / 8-11: row groups
constexpr int CH_SLOT = 128 * 32 * 4;      // one [128 x 32 floats] box = 16 KiB
constexpr int CH_MAX_RING = 8;
constexpr int CH_MAX_IN = 8;               // input boxes per tile
constexpr int CH_MAX_GEMM = 4;
constexpr uint32_t CH_SMEM_LIMIT = 227 * 1024;
// TMEM columns of one row group: [0,64) A hi, [64,128) A lo, [128,256) accumulator
constexpr int CH_TM_GROUP = 256, CH_TM_ALO = 64, CH_TM_D = 128;

struct ChainMaps {
  CUtensorMap in[4];
  CUtensorMap out[8];
  CUtensorMap w[2 * CH_MAX_GEMM];  // hi, lo image of each GEMM's B operand ([N, 64] row-major)
};

struct ChainCommon {
  int64_t R;
  int num_tiles;
  int n_in;                      // input boxes per tile, in the order the row threads consume them
  int in_map[CH_MAX_IN], in_col[CH_MAX_IN];
  int n_gemm;                    // GEMMs per tile, in issue order
  int g_N[CH_MAX_GEMM];
  int g_dcol[CH_MAX_GEMM];        // first accumulator column of the GEMM's result
  int g_wrow[CH_MAX_GEMM];       // first row of the GEMM's B operand inside its weight map
  uint32_t g_boff[CH_MAX_GEMM];  // byte offset of the hi image inside the weight area; the lo image follows it
  int ring_slots;
  int par_floats;                // size of the kernel's small-parameter area
  uint32_t w_bytes;              // whole weight area
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]   (kind::tf32, A-from-TMEM form: lane = row, one column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// ---- swizzled [128 x 32 floats] boxes: row r is 128 bytes, its 16-byte piece j sits at position j ^ (r & 7)
__device__ __forceinline__ void slot_read(const uint8_t* slot, int row, float (&v)[32]) {
  const float4* src = reinterpret_cast<const float4*>(slot + (size_t)row * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = src[j ^ (row & 7)];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
template <int OFF>
__device__ __forceinline__ void slot_write(uint8_t* slot, int row, const float (&v)[64]) {
  float4* dst = reinterpret_cast<float4*>(slot + (size_t)row * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j ^ (row & 7)] = make_float4(v[OFF + 4 * j], v[OFF + 4 * j + 1], v[OFF + 4 * j + 2], v[OFF + 4 * j + 3]);
}

// Everything one row thread needs to talk to the rest of the CTA.
struct RowCtx {
  int group, rt, lane;           // row group, row of the tile owned by this thread (= TMEM lane), lane in warp
  uint32_t tm;                   // TMEM address of this thread's lane, column 0 of its group
  uint8_t* ring;                 // input ring
  uint64_t *in_full, *in_empty;
  uint64_t *in_done_mine, *in_done_other;  // turn taking of the two row groups on the (in-order) input ring
  int ring_slots, n_in;
  uint32_t in_it;                // next input box (CTA-wide running index)
  uint8_t* obuf;                 // this group's 32 KiB output staging buffer
  uint64_t *a_ready, *d_ready;
  uint32_t a_uses, d_uses;
};

// The ring delivers the boxes of tile 0, 1, 2, ... in order and the two row groups take alternate tiles: a group may only start waiting
// for its tile's boxes once the other group has taken all boxes of the tile before (a parity wait cannot tell ring phases two apart).
__device__ __forceinline__ void in_begin(RowCtx& c, int ti) {
  c.in_it = (uint32_t)ti * (uint32_t)c.n_in;
  if (ti > 0) mbar_wait(c.in_done_other, (uint32_t)((ti - 1) >> 1) & 1u);
}
__device__ __forceinline__ void in_end(RowCtx& c) {
  __syncwarp();
  if (c.lane == 0) mbar_arrive(c.in_done_mine);
}
// next input box of this tile -> 32 floats of this thread's row
__device__ __forceinline__ void in_take(RowCtx& c, float (&v)[32]) {
  const uint32_t s = c.in_it % (uint32_t)c.ring_slots, ph = (c.in_it / (uint32_t)c.ring_slots) & 1u;
  mbar_wait(&c.in_full[s], ph);
  slot_read(c.ring + (size_t)s * CH_SLOT, c.rt, v);
  __syncwarp();
  if (c.lane == 0) mbar_arrive(&c.in_empty[s]);
  ++c.in_it;
}
__device__ __forceinline__ void in_take64(RowCtx& c, float (&x)[64]) {
  float v[32];
  in_take(c, v);
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = v[j];
  in_take(c, v);
#pragma unroll
  for (int j = 0; j < 32; ++j) x[32 + j] = v[j];
}

// this thread's 64-wide row -> A operand in tensor memory (hi = the TF32 truncation of each word, lo = the exact remainder)
__device__ __forceinline__ void put_A(RowCtx& c, const float (&x)[64]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t w = __float_as_uint(x[32 * h + j]) & 0xFFFFE000u;
      hi[j] = w;
      lo[j] = __float_as_uint(x[32 * h + j] - __uint_as_float(w));
    }
    tmem_st32(c.tm + 32 * h, hi);
    tmem_st32(c.tm + CH_TM_ALO + 32 * h, lo);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(c.a_ready);
  ++c.a_uses;
}
// the A operand stays as it is: hand it to the next GEMM of the program (after this thread has drained the accumulator it overwrites)
__device__ __forceinline__ void rearm_A(RowCtx& c) {
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(c.a_ready);
  ++c.a_uses;
}
__device__ __forceinline__ void wait_D(RowCtx& c) {
  mbar_wait(c.d_ready, c.d_uses & 1u);
  ++c.d_uses;
  tc_fence_after();
}
template <int COL>
__device__ __forceinline__ void ld_D64(const RowCtx& c, float (&x)[64]) {
  uint32_t v0[32], v1[32];
  tmem_ld32_nowait(c.tm + CH_TM_D + COL, v0);
  tmem_ld32_nowait(c.tm + CH_TM_D + COL + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    x[j] = __uint_as_float(v0[j]);
    x[32 + j] = __uint_as_float(v1[j]);
  }
}
__device__ __forceinline__ void ld_D32(const RowCtx& c, int col, float (&x)[32]) {
  uint32_t v[32];
  tmem_ld32(c.tm + CH_TM_D + col, v);
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
}

// the next two input boxes of this tile -> accumulator columns [COL, COL + 64) (parked until the chain needs them)
template <int COL>
__device__ __forceinline__ void stash_input64(RowCtx& c) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[32];
    in_take(c, v);
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v[j]);
    tmem_st32(c.tm + CH_TM_D + COL + 32 * h, w);
  }
  tmem_st_wait();
}
// x += accumulator columns [COL, COL + 64)
template <int COL>
__device__ __forceinline__ void add_D64(const RowCtx& c, float (&x)[64]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld32(c.tm + CH_TM_D + COL + 32 * h, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) x[32 * h + j] += __uint_as_float(v[j]);
  }
}

// this thread's 64-wide row -> columns [col0, col0 + 64) of an output tensor, rows [row0, row0 + 128). Every warp stages and stores its
// own 32 rows ([32 x 32 floats] TMA boxes out of the warp's 8 KiB of the group's buffer): no barrier between the warps of a group.
__device__ __forceinline__ void out_store64(RowCtx& c, const CUtensorMap* tm, int col0, int row0, const float (&x)[64]) {
  const int wrow = c.rt & ~31;
  uint8_t* buf = c.obuf + (size_t)wrow * 256;  // 2 x [32 rows x 128 B] per warp
  if (c.lane == 0) bulk_wait_read<0>();  // the stores that last used this buffer have finished reading it
  __syncwarp();
  slot_write<0>(buf, c.lane, x);
  slot_write<32>(buf + 32 * 128, c.lane, x);
  fence_proxy_async();
  __syncwarp();
  if (c.lane == 0) {
    tma_store_2d(tm, buf, col0, row0 + wrow);
    tma_store_2d(tm, buf + 32 * 128, col0 + 32, row0 + wrow);
    bulk_commit();
  }
}

}  // namespace chain
}  // namespace magpo
