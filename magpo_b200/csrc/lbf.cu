// LevelBasedForaging env + training wrapper stack, one warp per env, the env's entities staged in shared memory.
// Dynamics: jumanji 1.1.0 @ 9ced6b8 `environments/routing/lbf/{env,generator,observer,utils}.py` (third-party, not vendored in the
// reference; restated in oracle/lbf.py, which documents the algorithm and its known unknowns). Call sites in the reference:
// mava/utils/make_env.py:107-135 (construction), wrappers/jumanji.py:171-208 (LbfWrapper: float obs, team reward),
// wrappers/observation.py:42-54 (AgentIDWrapper), wrappers/auto_reset_wrapper.py:60-101, wrappers/episode_metrics.py:60-112.
#include <math_constants.h>

#include "common.cuh"
#include "envs.cuh"
#include "prng.cuh"

namespace magpo {

constexpr int kLbfMaxA = 8, kLbfMaxF = 8, kLbfMaxCells = 256, kLbfWarps = 4;
constexpr unsigned kFull = 0xffffffffu;

struct LbfEnv {  // one env, owned by one warp
  int apos[kLbfMaxA][2], alvl[kLbfMaxA], aload[kLbfMaxA];
  int fpos[kLbfMaxF][2], flvl[kLbfMaxF], feat[kLbfMaxF];
  int mv[kLbfMaxA][2], sum_adj[kLbfMaxF], eat_now[kLbfMaxF];
  int step;
  uint32_t key[2];
  float score[kLbfMaxCells];  // generator scratch: gumbel + log(mask)
  uint8_t cell[kLbfMaxCells]; // generator scratch: food placement mask
};

__device__ __forceinline__ int l1dist(const int* p, const int* q) { return abs(p[0] - q[0]) + abs(p[1] - q[1]); }

// RandomGenerator.__call__(key) for one env (generator.py), executed by a full warp.
__device__ void lbf_generate(const MagpoLbfCfg& c, LbfEnv& e, int lane, uint32_t k0, uint32_t k1) {
  const int G = c.grid_size, GG = G * G, A = c.num_agents, F = c.num_food;
  // key_food, key_agents, key_food_level, key_agent_level, key = split(key, 5)
  uint32_t s0 = 0, s1 = 0;
  if (lane < 5) prng_split_i(k0, k1, (uint32_t)lane, s0, s1);
  const uint32_t kf0 = __shfl_sync(kFull, s0, 0), kf1 = __shfl_sync(kFull, s1, 0);
  const uint32_t ka0 = __shfl_sync(kFull, s0, 1), ka1 = __shfl_sync(kFull, s1, 1);
  const uint32_t kl0 = __shfl_sync(kFull, s0, 2), kl1 = __shfl_sync(kFull, s1, 2);
  const uint32_t kg0 = __shfl_sync(kFull, s0, 3), kg1 = __shfl_sync(kFull, s1, 3);
  const uint32_t kn0 = __shfl_sync(kFull, s0, 4), kn1 = __shfl_sync(kFull, s1, 4);
  // sample_food: interior cells; every placed item removes its cell and the 4 neighbours
  for (int i = lane; i < GG; i += 32) {
    const int r = i / G, cc = i % G;
    e.cell[i] = (r > 0 && r < G - 1 && cc > 0 && cc < G - 1) ? 1 : 0;
  }
  __syncwarp();
  const int cpl = (GG + 31) / 32;
  const int lo = min(GG, lane * cpl), hi = min(GG, lo + cpl);
  for (int f = 0; f < F; ++f) {
    uint32_t fk0, fk1;
    prng_split_i(kf0, kf1, (uint32_t)f, fk0, fk1);
    // jax.random.choice(key, G*G, (), p=mask): r = cumsum[-1] * (1 - uniform(key, ())); searchsorted(cumsum, r) (left)
    const float u = __uint_as_float((prng_bits_i(fk0, fk1, 0) >> 9) | 0x3F800000u) - 1.0f;
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += e.cell[i];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    const float r = (float)total * (1.0f - u);
    const int k = (int)ceilf(r);  // the cumsum is integer valued: first index with cumsum >= r is the k-th open cell
    const int excl = incl - cnt;
    int pos = -1;
    if (k > excl && k <= incl) {
      int need = k - excl;
      for (int i = lo; i < hi; ++i)
        if (e.cell[i] && --need == 0) { pos = i; break; }
    }
    const unsigned m = __ballot_sync(kFull, pos >= 0);
    pos = __shfl_sync(kFull, pos, m ? __ffs(m) - 1 : 0);
    if (!m) pos = 0;
    __syncwarp();
    if (lane == 0) {
      e.fpos[f][0] = pos / G;
      e.fpos[f][1] = pos % G;
      const int adj[5] = {pos, pos + 1, pos - 1, pos + G, pos - G};
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (adj[j] >= 0 && adj[j] < GG) e.cell[adj[j]] = 0;
    }
    __syncwarp();
  }
  // sample_agents: choice(key, G*G, (A,), replace=False, p=mask) = top-A of gumbel + log(mask)
  for (int i = lane; i < GG; i += 32) {
    const int r = i / G, cc = i % G;
    bool ok = true;
    for (int f = 0; f < F; ++f) {
      if (c.agent_mask_rows) ok = ok && r != e.fpos[f][0] && r != e.fpos[f][1];
      else ok = ok && !(r == e.fpos[f][0] && cc == e.fpos[f][1]);
    }
    const float g = prng_gumbel_from_bits(prng_bits_i(ka0, ka1, (uint64_t)i));
    e.score[i] = ok ? g : -CUDART_INF_F;
  }
  __syncwarp();
  for (int a = 0; a < A; ++a) {
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int i = lane; i < GG; i += 32) {
      const float v = e.score[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(kFull, bv, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncwarp();
    if (lane == 0) {
      e.apos[a][0] = bi / G;
      e.apos[a][1] = bi % G;
      e.score[bi] = -CUDART_INF_F;
    }
    __syncwarp();
  }
  // levels
  uint32_t h0, h1, l0, l1;
  prng_split_i(kg0, kg1, 0u, h0, h1);
  prng_split_i(kg0, kg1, 1u, l0, l1);
  if (lane < A)
    e.alvl[lane] = prng_randint_from_bits(prng_bits_i(h0, h1, (uint64_t)lane), prng_bits_i(l0, l1, (uint64_t)lane), 1, c.max_agent_level + 1);
  __syncwarp();
  int m1 = 0x3fffffff, m2 = 0x3fffffff, m3 = 0x3fffffff;  // the three lowest agent levels
  for (int a = 0; a < A; ++a) {
    int v = e.alvl[a];
    if (v < m1) { m3 = m2; m2 = m1; m1 = v; }
    else if (v < m2) { m3 = m2; m2 = v; }
    else if (v < m3) { m3 = v; }
  }
  const int max_food_level = m1 + (A > 1 ? m2 : 0) + (A > 2 ? m3 : 0);
  if (lane < F) {
    int lvl = max_food_level;
    if (!c.force_coop) {
      prng_split_i(kl0, kl1, 0u, h0, h1);
      prng_split_i(kl0, kl1, 1u, l0, l1);
      lvl = prng_randint_from_bits(prng_bits_i(h0, h1, (uint64_t)lane), prng_bits_i(l0, l1, (uint64_t)lane), 1, max_food_level + 1);
    }
    e.flvl[lane] = lvl;
    e.feat[lane] = 0;
  }
  if (lane < A) e.aload[lane] = 0;
  if (lane == 0) {
    e.step = 0;
    e.key[0] = kn0;
    e.key[1] = kn1;
  }
  __syncwarp();
}

// VectorObserver.state_to_observation + compute_action_mask + AgentIDWrapper, written to up to two destinations.
__device__ void lbf_write_obs(const MagpoLbfCfg& c, const LbfEnv& e, int b, int lane, float* view, float* view2, uint8_t* mask,
                              int32_t* step_count, int32_t* step_count2) {
  const int G = c.grid_size, A = c.num_agents, F = c.num_food, fov = c.fov;
  const int d = A + 3 * (F + A);
  if (view || view2) {
    for (int i = lane; i < A * d; i += 32) {
      const int ag = i / d, f = i % d;
      float v;
      if (f < A) {
        v = f == ag ? 1.0f : 0.0f;
      } else {
        const int q = (f - A) / 3, comp = (f - A) % 3;
        const int* me = e.apos[ag];
        const int* p;
        int lvl;
        bool vis;
        if (q < F) {
          p = e.fpos[q];
          lvl = e.flvl[q];
          vis = !e.feat[q];
        } else {
          const int slot = q - F;
          const int j = slot == 0 ? ag : (slot - 1 < ag ? slot - 1 : slot);  // self first, then the others in id order
          p = e.apos[j];
          lvl = e.alvl[j];
          vis = true;
        }
        vis = vis && abs(me[0] - p[0]) <= fov && abs(me[1] - p[1]) <= fov;
        const int val = comp == 2 ? (vis ? lvl : 0) : (vis ? p[comp] - me[comp] + min(fov, me[comp]) : -1);
        v = (float)val;
      }
      if (view) view[(size_t)b * A * d + i] = v;
      if (view2) view2[(size_t)b * A * d + i] = v;
    }
  }
  if (mask) {
    for (int i = lane; i < A * 6; i += 32) {
      const int ag = i / 6, act = i % 6;
      const int dy = act == 1 ? -1 : (act == 2 ? 1 : 0), dx = act == 3 ? -1 : (act == 4 ? 1 : 0);
      const int ny = e.apos[ag][0] + dy, nx = e.apos[ag][1] + dx;
      bool ok = ny >= 0 && ny < G && nx >= 0 && nx < G;
      for (int j = 0; j < A; ++j) ok = ok && (j == ag || !(e.apos[j][0] == ny && e.apos[j][1] == nx));
      bool food_adj = false;
      for (int f = 0; f < F; ++f) {
        ok = ok && (e.feat[f] || !(e.fpos[f][0] == ny && e.fpos[f][1] == nx));
        food_adj = food_adj || (!e.feat[f] && l1dist(e.apos[ag], e.fpos[f]) == 1);
      }
      if (act == 5) ok = ok && food_adj;
      mask[(size_t)b * A * 6 + i] = ok ? 1 : 0;
    }
  }
  for (int i = lane; i < A; i += 32) {
    if (step_count) step_count[(size_t)b * A + i] = e.step;
    if (step_count2) step_count2[(size_t)b * A + i] = e.step;
  }
}

__device__ void lbf_store(const MagpoLbfCfg& c, const LbfEnv& e, const MagpoLbfState& st, int b, int lane) {
  const int A = c.num_agents, F = c.num_food;
  if (lane < A) {
    st.agent_pos[((size_t)b * A + lane) * 2] = e.apos[lane][0];
    st.agent_pos[((size_t)b * A + lane) * 2 + 1] = e.apos[lane][1];
    st.agent_level[(size_t)b * A + lane] = e.alvl[lane];
    st.agent_loading[(size_t)b * A + lane] = (uint8_t)e.aload[lane];
  }
  if (lane < F) {
    st.food_pos[((size_t)b * F + lane) * 2] = e.fpos[lane][0];
    st.food_pos[((size_t)b * F + lane) * 2 + 1] = e.fpos[lane][1];
    st.food_level[(size_t)b * F + lane] = e.flvl[lane];
    st.food_eaten[(size_t)b * F + lane] = (uint8_t)e.feat[lane];
  }
  if (lane == 0) {
    st.step_count[b] = e.step;
    st.key[2 * b] = e.key[0];
    st.key[2 * b + 1] = e.key[1];
  }
}

__global__ void lbf_reset_kernel(MagpoLbfCfg c, int B, const uint32_t* __restrict__ keys, MagpoLbfState st, MagpoTimeStep ts) {
  __shared__ LbfEnv envs[kLbfWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kLbfWarps + w;
  if (b >= B) return;
  LbfEnv& e = envs[w];
  const uint32_t k0 = keys[2 * b], k1 = keys[2 * b + 1];
  uint32_t m0, m1, r0, r1;
  prng_split_i(k0, k1, 0u, m0, m1);  // RecordEpisodeMetrics.reset: key, reset_key = split(key)
  prng_split_i(k0, k1, 1u, r0, r1);
  lbf_generate(c, e, lane, r0, r1);
  lbf_store(c, e, st, b, lane);
  lbf_write_obs(c, e, b, lane, ts.agents_view, ts.next_agents_view, ts.action_mask, ts.step_count, ts.next_step_count);
  const int A = c.num_agents;
  for (int i = lane; i < A; i += 32) {
    if (ts.reward) ts.reward[(size_t)b * A + i] = 0.0f;
    if (ts.discount) ts.discount[(size_t)b * A + i] = 1.0f;
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

__global__ void lbf_step_kernel(MagpoLbfCfg c, int B, const int32_t* __restrict__ action, MagpoLbfState st, MagpoTimeStep ts,
                                uint8_t* __restrict__ done_out) {
  __shared__ LbfEnv envs[kLbfWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kLbfWarps + w;
  if (b >= B) return;
  LbfEnv& e = envs[w];
  const int G = c.grid_size, A = c.num_agents, F = c.num_food;
  int act = 0;
  if (lane < A) {
    e.apos[lane][0] = st.agent_pos[((size_t)b * A + lane) * 2];
    e.apos[lane][1] = st.agent_pos[((size_t)b * A + lane) * 2 + 1];
    e.alvl[lane] = st.agent_level[(size_t)b * A + lane];
    act = min(max(action[(size_t)b * A + lane], 0), 5);
  }
  if (lane < F) {
    e.fpos[lane][0] = st.food_pos[((size_t)b * F + lane) * 2];
    e.fpos[lane][1] = st.food_pos[((size_t)b * F + lane) * 2 + 1];
    e.flvl[lane] = st.food_level[(size_t)b * F + lane];
    e.feat[lane] = st.food_eaten[(size_t)b * F + lane];
  }
  if (lane == 0) {
    e.step = st.step_count[b];
    e.key[0] = st.key[2 * b];
    e.key[1] = st.key[2 * b + 1];
  }
  __syncwarp();
  // update_agent_positions: simulate_agent_movement per agent, then one pass of fix_collisions
  if (lane < A) {
    const int dy = act == 1 ? -1 : (act == 2 ? 1 : 0), dx = act == 3 ? -1 : (act == 4 ? 1 : 0);
    const int ny = e.apos[lane][0] + dy, nx = e.apos[lane][1] + dx;
    bool blocked = ny < 0 || ny >= G || nx < 0 || nx >= G;
    for (int j = 0; j < A; ++j) blocked = blocked || (j != lane && e.apos[j][0] == ny && e.apos[j][1] == nx);
    for (int f = 0; f < F; ++f) blocked = blocked || (!e.feat[f] && e.fpos[f][0] == ny && e.fpos[f][1] == nx);
    e.mv[lane][0] = blocked ? e.apos[lane][0] : ny;
    e.mv[lane][1] = blocked ? e.apos[lane][1] : nx;
  }
  __syncwarp();
  int py = 0, px = 0;
  if (lane < A) {
    bool dup = false;
    for (int j = 0; j < A; ++j) dup = dup || (j != lane && e.mv[j][0] == e.mv[lane][0] && e.mv[j][1] == e.mv[lane][1]);
    py = dup ? e.apos[lane][0] : e.mv[lane][0];
    px = dup ? e.apos[lane][1] : e.mv[lane][1];
  }
  __syncwarp();
  if (lane < A) {
    e.apos[lane][0] = py;
    e.apos[lane][1] = px;
    e.aload[lane] = act == 5;
  }
  __syncwarp();
  // eat_food per item
  if (lane < F) {
    int s = 0;
    for (int a = 0; a < A; ++a) s += (l1dist(e.apos[a], e.fpos[lane]) == 1 && e.aload[a] && !e.feat[lane]) ? e.alvl[a] : 0;
    e.sum_adj[lane] = s;
    e.eat_now[lane] = s >= e.flvl[lane];
  }
  __syncwarp();
  // get_reward (normalised) -> LbfWrapper team reward -> RecordEpisodeMetrics mean; every lane computes the same scalars
  int total_food_level = 0;
  for (int f = 0; f < F; ++f) total_food_level += e.flvl[f];
  float team = 0.0f;
  for (int a = 0; a < A; ++a) {
    float r = 0.0f;
    for (int f = 0; f < F; ++f) {
      const int adjl = (l1dist(e.apos[a], e.fpos[f]) == 1 && e.aload[a] && !e.feat[f]) ? e.alvl[a] : 0;
      const int num = adjl * e.eat_now[f] * e.flvl[f];
      const int den = e.sum_adj[f] * total_food_level;
      r += den ? __fdiv_rn((float)num, (float)den) : 0.0f;  // nan_to_num(0/0) = 0
    }
    team = a == 0 ? r : team + r;
  }
  float mean_r = team;
  for (int a = 1; a < A; ++a) mean_r += team;
  mean_r = __fdiv_rn(mean_r, (float)A);
  __syncwarp();
  bool all_eaten = true;
  for (int f = 0; f < F; ++f) all_eaten = all_eaten && (e.feat[f] || e.eat_now[f]);
  __syncwarp();
  if (lane < F) e.feat[lane] = e.feat[lane] || e.eat_now[lane];
  const int steps = e.step + 1;
  __syncwarp();
  if (lane == 0) e.step = steps;
  __syncwarp();
  const bool terminate = all_eaten, truncate = steps >= c.time_limit;
  const bool done = terminate || truncate;
  if (done) {
    lbf_write_obs(c, e, b, lane, ts.next_agents_view, nullptr, nullptr, ts.next_step_count, nullptr);  // real_next_obs
    __syncwarp();
    uint32_t r0, r1;
    prng_split_i(e.key[0], e.key[1], 0u, r0, r1);  // key, _ = split(state.key)
    __syncwarp();
    lbf_generate(c, e, lane, r0, r1);
    lbf_write_obs(c, e, b, lane, ts.agents_view, nullptr, ts.action_mask, ts.step_count, nullptr);
  } else {
    lbf_write_obs(c, e, b, lane, ts.agents_view, ts.next_agents_view, ts.action_mask, ts.step_count, ts.next_step_count);
  }
  lbf_store(c, e, st, b, lane);
  for (int i = lane; i < A; i += 32) {
    if (ts.reward) ts.reward[(size_t)b * A + i] = team;
    if (ts.discount) ts.discount[(size_t)b * A + i] = terminate ? 0.0f : 1.0f;
  }
  if (lane == 0) {
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

static int lbf_check(const MagpoLbfCfg* c) {
  if (c->grid_size < 3 || c->num_agents < 1 || c->num_food < 1 || c->max_agent_level < 1 || c->time_limit < 1 || c->fov < 0)
    return MAGPO_ERR_ARG;
  if (c->grid_size * c->grid_size > kLbfMaxCells || c->num_agents > kLbfMaxA || c->num_food > kLbfMaxF) return MAGPO_ERR_UNSUPPORTED;
  return MAGPO_OK;
}

int lbf_step_launch(cudaStream_t s, const MagpoLbfCfg* cfg, int B, const int32_t* action, MagpoLbfState st, MagpoTimeStep ts,
                    uint8_t* done_out) {
  MAGPO_TRY(lbf_check(cfg));
  const int A = cfg->num_agents, F = cfg->num_food, d = A + 3 * (F + A);
  // algorithmic bytes per env-step: state in + out (13A + 13F + 12 + 16 B each way), actions in; obs x2, mask, step counts, reward,
  // discount, metrics out
  ProfScope ps(PROF_ENV, s, (double)B * (2.0 * (13 * A + 13 * F + 28) + 4 * A + 2 * 4 * A * d + 6 * A + 8 * A + 8 * A + 12));
  lbf_step_kernel<<<(unsigned)ceil_div(B, kLbfWarps), kLbfWarps * 32, 0, s>>>(*cfg, B, action, st, ts, done_out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo

using namespace magpo;

extern "C" {

int magpo_lbf_reset(magpo_stream_t s, const MagpoLbfCfg* cfg, int32_t B, const uint32_t* keys, MagpoLbfState st, MagpoTimeStep ts) {
  if (!cfg || !keys || B < 0) return MAGPO_ERR_ARG;
  MAGPO_TRY(lbf_check(cfg));
  if (B == 0) return MAGPO_OK;
  lbf_reset_kernel<<<(unsigned)ceil_div(B, kLbfWarps), kLbfWarps * 32, 0, as_stream(s)>>>(*cfg, B, keys, st, ts);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_lbf_step(magpo_stream_t s, const MagpoLbfCfg* cfg, int32_t B, const int32_t* action, MagpoLbfState st, MagpoTimeStep ts) {
  if (!cfg || !action || B < 0) return MAGPO_ERR_ARG;
  if (B == 0) return MAGPO_OK;
  return lbf_step_launch(as_stream(s), cfg, B, action, st, ts, nullptr);
}

}  // extern "C"
