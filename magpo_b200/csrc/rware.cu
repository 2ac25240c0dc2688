// RobotWarehouse env + training wrapper stack, one warp per env, the two grid layers staged in shared memory.
// Dynamics: jumanji 1.1.0 @ 9ced6b8 `environments/routing/robot_warehouse/*` (third-party, not vendored in the reference; restated in
// oracle/rware.py, which documents the algorithm and its known unknowns). Call sites in the reference:
// mava/utils/make_env.py:107-135 (construction), wrappers/jumanji.py:137-168 (RwareWrapper: float obs, reward/discount repeated),
// wrappers/observation.py:42-54 (AgentIDWrapper), wrappers/auto_reset_wrapper.py:60-101, wrappers/episode_metrics.py:60-112.
#include <math_constants.h>

#include "common.cuh"
#include "envs.cuh"
#include "prng.cuh"

namespace magpo {

constexpr int kRwMaxA = 8, kRwMaxQ = 16, kRwMaxCells = 1024, kRwWarps = 4;
constexpr unsigned kFullMask = 0xffffffffu;

struct RwDims {
  int H, W, cells, S, A, Q, sr, ch, time_limit;
};

struct RwSmall {
  int apos[kRwMaxA][2], adir[kRwMaxA], acarry[kRwMaxA], queue[kRwMaxQ];
  int step;
  uint32_t key[2];
  uint8_t amask[kRwMaxA * 5];
};

struct RwEnv {  // views into one warp's slice of dynamic shared memory
  int32_t* g0;   // [cells] shelf id + 1
  int32_t* g1;   // [cells] agent id + 1
  uint8_t* req;  // [S]
  RwSmall* s;
};

__host__ __device__ __forceinline__ bool rw_highway(const RwDims& d, int r, int c) {
  return (c % 3 == 0) || (r % (d.ch + 1) == 0) || (r == d.H - 1) || (r > d.H - (d.ch + 3) && (c == d.W / 2 - 1 || c == d.W / 2));
}

static RwDims rw_dims(const MagpoRwareCfg* c) {
  RwDims d;
  d.ch = c->column_height;
  d.H = (c->column_height + 1) * c->shelf_rows + 2;
  d.W = 3 * c->shelf_columns + 1;
  d.cells = d.H * d.W;
  d.A = c->num_agents;
  d.Q = c->request_queue_size;
  d.sr = c->sensor_range;
  d.time_limit = c->time_limit;
  d.S = 0;
  for (int r = 0; r < d.H; ++r)
    for (int cc = 0; cc < d.W; ++cc) d.S += rw_highway(d, r, cc) ? 0 : 1;
  return d;
}

__host__ __device__ __forceinline__ size_t rw_warp_bytes(const RwDims& d) {
  return (size_t)d.cells * 8 + (((size_t)d.S + 15) & ~size_t(15)) + ((sizeof(RwSmall) + 15) & ~size_t(15));
}

__device__ __forceinline__ RwEnv rw_env(const RwDims& d, unsigned char* smem, int w) {
  unsigned char* p = smem + (size_t)w * rw_warp_bytes(d);
  RwEnv e;
  e.g0 = reinterpret_cast<int32_t*>(p);
  e.g1 = e.g0 + d.cells;
  e.req = p + (size_t)d.cells * 8;
  e.s = reinterpret_cast<RwSmall*>(e.req + (((size_t)d.S + 15) & ~size_t(15)));
  return e;
}

// First `k` entries of jax.random.permutation(key, n) for n <= 1625 (one sort round): `key, sub = split(key)`, stable sort of
// arange(n) by random_bits(sub, (n,)) -> the k smallest (bits, index) pairs in order. `scratch` holds n words. out: lane 0 writes.
__device__ void rw_perm_head(uint32_t k0, uint32_t k1, int n, int k, uint32_t* scratch, int* out, int lane) {
  uint32_t s0, s1;
  prng_split_i(k0, k1, 1u, s0, s1);
  for (int i = lane; i < n; i += 32) scratch[i] = prng_bits_i(s0, s1, (uint64_t)i);
  __syncwarp();
  uint64_t prev = 0;
  bool have_prev = false;
  for (int j = 0; j < k; ++j) {
    uint64_t best = ~0ull;
    for (int i = lane; i < n; i += 32) {
      const uint64_t v = ((uint64_t)scratch[i] << 32) | (uint32_t)i;
      if ((!have_prev || v > prev) && v < best) best = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t ov = __shfl_xor_sync(kFullMask, best, o);
      best = ov < best ? ov : best;
    }
    if (lane == 0) out[j] = (int)(uint32_t)best;
    prev = best;
    have_prev = true;
  }
  __syncwarp();
}

__device__ void rw_compute_mask(const RwDims& d, RwEnv& e, int lane) {
  if (lane < d.A) {
    const int dir = e.s->adir[lane];
    const int nx = e.s->apos[lane][0] + (dir == 0 ? -1 : (dir == 2 ? 1 : 0));
    const int ny = e.s->apos[lane][1] + (dir == 1 ? 1 : (dir == 3 ? -1 : 0));
    bool ok = nx >= 0 && nx < d.H && ny >= 0 && ny < d.W;
    if (ok) ok = e.g1[nx * d.W + ny] == 0 && !(e.s->acarry[lane] && e.g0[nx * d.W + ny] > 0);
#pragma unroll
    for (int a = 0; a < 5; ++a) e.s->amask[lane * 5 + a] = (a == 1) ? (ok ? 1 : 0) : 1;
  }
  __syncwarp();
}

// RandomGenerator.__call__(key) for one env, executed by a full warp; also writes the (static) shelf positions.
__device__ void rw_generate(const RwDims& d, RwEnv& e, int lane, uint32_t k0, uint32_t k1, int32_t* shelf_pos_out) {
  uint32_t s0 = 0, s1 = 0;
  if (lane < 4) prng_split_i(k0, k1, (uint32_t)lane, s0, s1);  // key, agent_key, dir_key, queue_key = split(key, 4)
  const uint32_t kn0 = __shfl_sync(kFullMask, s0, 0), kn1 = __shfl_sync(kFullMask, s1, 0);
  const uint32_t ka0 = __shfl_sync(kFullMask, s0, 1), ka1 = __shfl_sync(kFullMask, s1, 1);
  const uint32_t kd0 = __shfl_sync(kFullMask, s0, 2), kd1 = __shfl_sync(kFullMask, s1, 2);
  const uint32_t kq0 = __shfl_sync(kFullMask, s0, 3), kq1 = __shfl_sync(kFullMask, s1, 3);
  __shared__ int picks[kRwWarps][kRwMaxQ > kRwMaxA ? kRwMaxQ : kRwMaxA];
  int* pick = picks[threadIdx.x >> 5];
  uint32_t* scratch = reinterpret_cast<uint32_t*>(e.g0);  // the grid layers are rebuilt below
  rw_perm_head(ka0, ka1, d.cells, d.A, scratch, pick, lane);
  if (lane < d.A) {
    e.s->apos[lane][0] = pick[lane] / d.W;
    e.s->apos[lane][1] = pick[lane] % d.W;
    uint32_t h0, h1, l0, l1;
    prng_split_i(kd0, kd1, 0u, h0, h1);
    prng_split_i(kd0, kd1, 1u, l0, l1);
    e.s->adir[lane] = prng_randint_from_bits(prng_bits_i(h0, h1, (uint64_t)lane), prng_bits_i(l0, l1, (uint64_t)lane), 0, 4);
    e.s->acarry[lane] = 0;
  }
  __syncwarp();
  rw_perm_head(kq0, kq1, d.S, d.Q, scratch, pick, lane);
  if (lane < d.Q) e.s->queue[lane] = pick[lane];
  for (int i = lane; i < d.S; i += 32) e.req[i] = 0;
  __syncwarp();
  if (lane < d.Q) e.req[pick[lane]] = 1;
  // grid: shelves on every non-highway cell, ids in row-major order
  const int cpl = (d.cells + 31) / 32;
  const int lo = min(d.cells, lane * cpl), hi = min(d.cells, lo + cpl);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += rw_highway(d, i / d.W, i % d.W) ? 0 : 1;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += v;
  }
  int id = incl - cnt;
  __syncwarp();
  for (int i = lo; i < hi; ++i) {
    const bool shelf = !rw_highway(d, i / d.W, i % d.W);
    e.g0[i] = shelf ? id + 1 : 0;
    e.g1[i] = 0;
    if (shelf) {
      shelf_pos_out[2 * id] = i / d.W;
      shelf_pos_out[2 * id + 1] = i % d.W;
      ++id;
    }
  }
  __syncwarp();
  if (lane < d.A) e.g1[e.s->apos[lane][0] * d.W + e.s->apos[lane][1]] = lane + 1;
  if (lane == 0) {
    e.s->step = 0;
    e.s->key[0] = kn0;
    e.s->key[1] = kn1;
  }
  __syncwarp();
  rw_compute_mask(d, e, lane);
}

// make_agent_observation for every agent + AgentIDWrapper, written to up to two destinations; action mask from the state.
__device__ void rw_write_obs(const RwDims& d, const RwEnv& e, int b, int lane, float* view, float* view2, uint8_t* mask,
                             int32_t* step_count, int32_t* step_count2) {
  const int A = d.A, rf = 2 * d.sr + 1;
  const int od = A + 8 + 7 * rf * rf;
  if (view || view2) {
    for (int i = lane; i < A * od; i += 32) {
      const int ag = i / od, f = i % od;
      const int x = e.s->apos[ag][0], y = e.s->apos[ag][1];
      int val;
      if (f < A) {
        val = f == ag;
      } else {
        const int v = f - A;
        if (v == 0) val = x;
        else if (v == 1) val = y;
        else if (v == 2) val = e.s->acarry[ag];
        else if (v < 7) val = e.s->adir[ag] == v - 3;
        else if (v == 7) val = rw_highway(d, x, y);
        else {
          const int cell = (v - 8) / 7, comp = (v - 8) % 7;
          const int gx = x + cell / rf - d.sr, gy = y + cell % rf - d.sr;
          int aid = 0, sid = 0;
          if (gx >= 0 && gx < d.H && gy >= 0 && gy < d.W) {
            aid = e.g1[gx * d.W + gy];
            sid = e.g0[gx * d.W + gy];
          }
          if (comp == 0) val = aid > 0;
          else if (comp < 5) val = aid > 0 && e.s->adir[aid - 1] == comp - 1;
          else if (comp == 5) val = sid > 0;
          else val = sid > 0 && e.req[sid - 1];
        }
      }
      if (view) view[(size_t)b * A * od + i] = (float)val;
      if (view2) view2[(size_t)b * A * od + i] = (float)val;
    }
  }
  if (mask)
    for (int i = lane; i < A * 5; i += 32) mask[(size_t)b * A * 5 + i] = e.s->amask[i];
  for (int i = lane; i < A; i += 32) {
    if (step_count) step_count[(size_t)b * A + i] = e.s->step;
    if (step_count2) step_count2[(size_t)b * A + i] = e.s->step;
  }
}

__device__ void rw_store(const RwDims& d, const RwEnv& e, const MagpoRwareState& st, int b, int lane) {
  int32_t* g = st.grid + (size_t)b * 2 * d.cells;
  for (int i = lane; i < 2 * d.cells; i += 32) g[i] = e.g0[i];  // g1 follows g0 in shared memory
  for (int i = lane; i < d.S; i += 32) st.shelf_req[(size_t)b * d.S + i] = e.req[i];
  if (lane < d.A) {
    st.agent_pos[((size_t)b * d.A + lane) * 2] = e.s->apos[lane][0];
    st.agent_pos[((size_t)b * d.A + lane) * 2 + 1] = e.s->apos[lane][1];
    st.agent_dir[(size_t)b * d.A + lane] = e.s->adir[lane];
    st.agent_carry[(size_t)b * d.A + lane] = (uint8_t)e.s->acarry[lane];
  }
  for (int i = lane; i < d.A * 5; i += 32) st.action_mask[(size_t)b * d.A * 5 + i] = e.s->amask[i];
  if (lane < d.Q) st.request_queue[(size_t)b * d.Q + lane] = e.s->queue[lane];
  if (lane == 0) {
    st.step_count[b] = e.s->step;
    st.key[2 * b] = e.s->key[0];
    st.key[2 * b + 1] = e.s->key[1];
  }
}

__global__ void rware_reset_kernel(RwDims d, int B, const uint32_t* __restrict__ keys, MagpoRwareState st, MagpoTimeStep ts) {
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kRwWarps + w;
  if (b >= B) return;
  RwEnv e = rw_env(d, rw_smem, w);
  const uint32_t k0 = keys[2 * b], k1 = keys[2 * b + 1];
  uint32_t m0, m1, r0, r1;
  prng_split_i(k0, k1, 0u, m0, m1);  // RecordEpisodeMetrics.reset: key, reset_key = split(key)
  prng_split_i(k0, k1, 1u, r0, r1);
  rw_generate(d, e, lane, r0, r1, st.shelf_pos + (size_t)b * d.S * 2);
  rw_store(d, e, st, b, lane);
  rw_write_obs(d, e, b, lane, ts.agents_view, ts.next_agents_view, ts.action_mask, ts.step_count, ts.next_step_count);
  for (int i = lane; i < d.A; i += 32) {
    if (ts.reward) ts.reward[(size_t)b * d.A + i] = 0.0f;
    if (ts.discount) ts.discount[(size_t)b * d.A + i] = 1.0f;
  }
  if (lane == 0) {
    st.metrics_key[2 * b] = m0;
    st.metrics_key[2 * b + 1] = m1;
    st.running_return[b] = 0.0f;
    st.running_length[b] = 0;
    st.episode_return[b] = 0.0f;
    st.episode_length[b] = 0;
    if (ts.step_type) ts.step_type[b] = 0;
    if (ts.episode_return) ts.episode_return[b] = 0.0f;
    if (ts.episode_length) ts.episode_length[b] = 0;
    if (ts.is_terminal_step) ts.is_terminal_step[b] = 0;
  }
}

__global__ void rware_step_kernel(RwDims d, int B, const int32_t* __restrict__ action, MagpoRwareState st, MagpoTimeStep ts,
                                  uint8_t* __restrict__ done_out) {
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kRwWarps + w;
  if (b >= B) return;
  RwEnv e = rw_env(d, rw_smem, w);
  const int A = d.A, W = d.W;
  {  // stage the env
    const int32_t* g = st.grid + (size_t)b * 2 * d.cells;
    for (int i = lane; i < 2 * d.cells; i += 32) e.g0[i] = g[i];
    for (int i = lane; i < d.S; i += 32) e.req[i] = st.shelf_req[(size_t)b * d.S + i];
    if (lane < A) {
      e.s->apos[lane][0] = st.agent_pos[((size_t)b * A + lane) * 2];
      e.s->apos[lane][1] = st.agent_pos[((size_t)b * A + lane) * 2 + 1];
      e.s->adir[lane] = st.agent_dir[(size_t)b * A + lane];
      e.s->acarry[lane] = st.agent_carry[(size_t)b * A + lane];
    }
    for (int i = lane; i < A * 5; i += 32) e.s->amask[i] = st.action_mask[(size_t)b * A * 5 + i];
    if (lane < d.Q) e.s->queue[lane] = st.request_queue[(size_t)b * d.Q + lane];
    if (lane == 0) {
      e.s->step = st.step_count[b];
      e.s->key[0] = st.key[2 * b];
      e.s->key[1] = st.key[2 * b + 1];
    }
  }
  __syncwarp();
  if (lane == 0) {  // agents are updated one after the other on the shared grid (lax.scan over agents)
    int32_t* spos = st.shelf_pos + (size_t)b * d.S * 2;
    for (int i = 0; i < A; ++i) {
      int act = action[(size_t)b * A + i];
      if (act < 0 || act > 4 || !e.s->amask[i * 5 + act]) act = 0;
      const int x = e.s->apos[i][0], y = e.s->apos[i][1], dir = e.s->adir[i];
      if (act == 2) e.s->adir[i] = (dir + 3) & 3;
      else if (act == 3) e.s->adir[i] = (dir + 1) & 3;
      else if (act == 1) {
        const int nx = min(max(x + (dir == 0 ? -1 : (dir == 2 ? 1 : 0)), 0), d.H - 1);
        const int ny = min(max(y + (dir == 1 ? 1 : (dir == 3 ? -1 : 0)), 0), W - 1);
        e.g1[x * W + y] = 0;
        e.g1[nx * W + ny] = i + 1;
        if (e.s->acarry[i]) {
          const int sid = e.g0[x * W + y];
          e.g0[x * W + y] = 0;
          e.g0[nx * W + ny] = sid;
          if (sid > 0) {
            spos[2 * (sid - 1)] = nx;
            spos[2 * (sid - 1) + 1] = ny;
          }
        }
        e.s->apos[i][0] = nx;
        e.s->apos[i][1] = ny;
      } else if (act == 4) {
        if (!e.s->acarry[i]) e.s->acarry[i] = e.g0[x * W + y] > 0;
        else if (!rw_highway(d, x, y)) e.s->acarry[i] = 0;
      }
    }
  }
  __syncwarp();
  const bool my_coll = lane < A && e.g1[e.s->apos[lane][0] * W + e.s->apos[lane][1]] != lane + 1;
  const bool collision = __ballot_sync(kFullMask, my_coll) != 0;
  // goals, in order: a requested shelf on a goal cell is delivered and a new request drawn
  float reward = 0.0f;
  uint32_t key0 = e.s->key[0], key1 = e.s->key[1];
  for (int gidx = 0; gidx < 2; ++gidx) {
    const int gcell = (d.H - 1) * W + (W / 2 - 1 + gidx);
    const int sid = e.g0[gcell];
    if (sid > 0 && e.req[sid - 1]) {  // warp-uniform
      reward += 1.0f;
      uint32_t n0, n1, r0, r1;
      prng_split_i(key0, key1, 0u, n0, n1);  // key, request_key = split(key)
      prng_split_i(key0, key1, 1u, r0, r1);
      key0 = n0;
      key1 = n1;
      float bv = -CUDART_INF_F;
      int bi = 0x7fffffff;
      for (int i = lane; i < d.S; i += 32) {
        const float v = e.req[i] ? -CUDART_INF_F : prng_gumbel_from_bits(prng_bits_i(r0, r1, (uint64_t)i));
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(kFullMask, bv, o);
        const int oi = __shfl_xor_sync(kFullMask, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      __syncwarp();
      if (lane < d.Q && e.s->queue[lane] == sid - 1) e.s->queue[lane] = bi;
      if (lane == 0) {
        e.req[sid - 1] = 0;
        e.req[bi] = 1;
      }
      __syncwarp();
    }
  }
  const int steps = e.s->step + 1;
  __syncwarp();
  if (lane == 0) {
    e.s->step = steps;
    e.s->key[0] = key0;
    e.s->key[1] = key1;
  }
  __syncwarp();
  rw_compute_mask(d, e, lane);
  const bool done = collision || steps >= d.time_limit;
  if (done) {
    rw_write_obs(d, e, b, lane, ts.next_agents_view, nullptr, nullptr, ts.next_step_count, nullptr);  // real_next_obs
    __syncwarp();
    uint32_t r0, r1;
    prng_split_i(key0, key1, 0u, r0, r1);  // key, _ = split(state.key)
    rw_generate(d, e, lane, r0, r1, st.shelf_pos + (size_t)b * d.S * 2);
    rw_write_obs(d, e, b, lane, ts.agents_view, nullptr, ts.action_mask, ts.step_count, nullptr);
  } else {
    rw_write_obs(d, e, b, lane, ts.agents_view, ts.next_agents_view, ts.action_mask, ts.step_count, ts.next_step_count);
  }
  rw_store(d, e, st, b, lane);
  for (int i = lane; i < A; i += 32) {
    if (ts.reward) ts.reward[(size_t)b * A + i] = reward;
    if (ts.discount) ts.discount[(size_t)b * A + i] = done ? 0.0f : 1.0f;
  }
  if (lane == 0) {
    float mean_r = reward;  // RecordEpisodeMetrics: mean of A identical rewards = (r + ... + r) / A
    for (int a = 1; a < A; ++a) mean_r += reward;
    mean_r = __fdiv_rn(mean_r, (float)A);
    const float new_ret = st.running_return[b] + mean_r;
    const int new_len = st.running_length[b] + 1;
    const float ep_ret = done ? new_ret : st.episode_return[b];
    const int ep_len = done ? new_len : st.episode_length[b];
    st.running_return[b] = done ? 0.0f : new_ret;
    st.running_length[b] = done ? 0 : new_len;
    st.episode_return[b] = ep_ret;
    st.episode_length[b] = ep_len;
    if (ts.step_type) ts.step_type[b] = done ? 2 : 1;
    if (ts.episode_return) ts.episode_return[b] = ep_ret;
    if (ts.episode_length) ts.episode_length[b] = ep_len;
    if (ts.is_terminal_step) ts.is_terminal_step[b] = done ? 1 : 0;
    if (done_out) done_out[b] = done ? 1 : 0;
  }
}

static int rw_check(const MagpoRwareCfg* c, RwDims* out) {
  if (c->column_height < 1 || c->shelf_rows < 1 || c->shelf_columns < 1 || c->num_agents < 1 || c->sensor_range < 0 ||
      c->request_queue_size < 1 || c->time_limit < 1)
    return MAGPO_ERR_ARG;
  RwDims d = rw_dims(c);
  if (d.cells > kRwMaxCells || d.A > kRwMaxA || d.Q > kRwMaxQ || d.sr > 2 || d.Q > d.S || d.A > d.cells) return MAGPO_ERR_UNSUPPORTED;
  *out = d;
  return MAGPO_OK;
}

int rware_step_launch(cudaStream_t s, const MagpoRwareCfg* cfg, int B, const int32_t* action, MagpoRwareState st, MagpoTimeStep ts,
                      uint8_t* done_out) {
  RwDims d;
  MAGPO_TRY(rw_check(cfg, &d));
  const size_t smem = kRwWarps * rw_warp_bytes(d);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MAGPO_CUDA_OK(cudaFuncSetAttribute(rware_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int od = d.A + 8 + 7 * (2 * d.sr + 1) * (2 * d.sr + 1);
  // algorithmic bytes per env-step: state in + out (grid 8 cells, requests S, agents 13 A, mask 5 A, queue 4 Q, 12), actions in;
  // obs x2, mask, step counts, reward, discount, metrics out
  ProfScope ps(PROF_ENV, s, (double)B * (2.0 * (8.0 * d.cells + d.S + 18 * d.A + 4 * d.Q + 28) + 4 * d.A + 2 * 4.0 * d.A * od +
                                          5 * d.A + 8 * d.A + 8 * d.A + 12));
  rware_step_kernel<<<(unsigned)ceil_div(B, kRwWarps), kRwWarps * 32, smem, s>>>(d, B, action, st, ts, done_out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo

using namespace magpo;

extern "C" {

int32_t magpo_rware_num_shelves(const MagpoRwareCfg* cfg) {
  if (!cfg) return MAGPO_ERR_ARG;
  RwDims d;
  MAGPO_TRY(rw_check(cfg, &d));
  return d.S;
}

int magpo_rware_reset(magpo_stream_t s, const MagpoRwareCfg* cfg, int32_t B, const uint32_t* keys, MagpoRwareState st,
                      MagpoTimeStep ts) {
  if (!cfg || !keys || B < 0) return MAGPO_ERR_ARG;
  RwDims d;
  MAGPO_TRY(rw_check(cfg, &d));
  if (B == 0) return MAGPO_OK;
  const size_t smem = kRwWarps * rw_warp_bytes(d);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MAGPO_CUDA_OK(cudaFuncSetAttribute(rware_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  rware_reset_kernel<<<(unsigned)ceil_div(B, kRwWarps), kRwWarps * 32, smem, as_stream(s)>>>(d, B, keys, st, ts);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_rware_step(magpo_stream_t s, const MagpoRwareCfg* cfg, int32_t B, const int32_t* action, MagpoRwareState st,
                     MagpoTimeStep ts) {
  if (!cfg || !action || B < 0) return MAGPO_ERR_ARG;
  if (B == 0) return MAGPO_OK;
  return rware_step_launch(as_stream(s), cfg, B, action, st, ts, nullptr);
}

}  // extern "C"
