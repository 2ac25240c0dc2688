// CoordSum env + training wrapper stack, one warp per env.
// Reference: mava/coordsum/env.py:55-139 (reset/step), wrappers/matrax.py:117-134 (CoordSumWrapper),
// wrappers/observation.py:42-54 (AgentIDWrapper), wrappers/auto_reset_wrapper.py:60-101,
// wrappers/episode_metrics.py:60-112.  Quirks kept: record-row index clamps to num_actions-1
// (JAX gather / dynamic_update_slice clamp), bincount(length=time_limit) first-max argmax,
// reward 1 when the modal past first-action equals actions[0] else 2, auto-reset key = split(state.key)[0].
#include "common.cuh"
#include "envs.cuh"
#include "prng.cuh"

namespace magpo {

// CoordSum.reset(key) for one env, executed by a full warp.
__device__ void coordsum_base_reset(const MagpoCoordSumCfg& c, const MagpoCoordSumState& st, int b, int lane,
                                    uint32_t k0, uint32_t k1) {
  uint32_t nk0, nk1, t0, t1;
  prng_split_i(k0, k1, 0u, nk0, nk1);  // key, target_key = split(key)
  prng_split_i(k0, k1, 1u, t0, t1);
  uint32_t a0, a1, b0, b1;
  prng_split_i(t0, t1, 0u, a0, a1);  // randint: k1, k2 = split(target_key)
  prng_split_i(t0, t1, 1u, b0, b1);
  const int TL = c.time_limit;
  int32_t* target = st.target + (size_t)b * (TL + 1);
  for (int i = lane; i <= TL; i += 32) {
    uint32_t hi = prng_bits_i(a0, a1, (uint64_t)i);
    uint32_t lo = prng_bits_i(b0, b1, (uint64_t)i);
    target[i] = prng_randint_from_bits(hi, lo, 0, c.maxval);
  }
  int32_t* rec = st.record + (size_t)b * c.num_actions * TL;
  for (int i = lane; i < c.num_actions * TL; i += 32) rec[i] = -1;
  if (lane == 0) {
    st.step_count[b] = 0;
    st.key[2 * b] = nk0;
    st.key[2 * b + 1] = nk1;
  }
}

__device__ void coordsum_write_obs(const MagpoCoordSumCfg& c, int b, int lane, int32_t target_val, int32_t step,
                                   float* view, uint8_t* mask, int32_t* step_count) {
  const int A = c.num_agents, d = A + 1, a = c.num_actions;
  if (view) {
    float* v = view + (size_t)b * A * d;
    for (int i = lane; i < A * d; i += 32) {
      int ag = i / d, f = i % d;
      v[i] = f == A ? (float)target_val : (f == ag ? 1.0f : 0.0f);
    }
  }
  if (mask) {
    uint8_t* m = mask + (size_t)b * A * a;
    for (int i = lane; i < A * a; i += 32) m[i] = 1;
  }
  for (int i = lane; step_count && i < A; i += 32) step_count[(size_t)b * A + i] = step;
}

__global__ void coordsum_reset_kernel(MagpoCoordSumCfg c, int B, const uint32_t* __restrict__ keys,
                                      MagpoCoordSumState st, MagpoTimeStep ts) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  uint32_t k0 = keys[2 * b], k1 = keys[2 * b + 1];
  uint32_t m0, m1, r0, r1;
  prng_split_i(k0, k1, 0u, m0, m1);  // RecordEpisodeMetrics.reset: key, reset_key = split(key)
  prng_split_i(k0, k1, 1u, r0, r1);
  coordsum_base_reset(c, st, b, lane, r0, r1);
  __syncwarp();
  int32_t tv = st.target[(size_t)b * (c.time_limit + 1)];
  coordsum_write_obs(c, b, lane, tv, 0, ts.agents_view, ts.action_mask, ts.step_count);
  coordsum_write_obs(c, b, lane, tv, 0, ts.next_agents_view, nullptr, ts.next_step_count);
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

constexpr int kStepWarps = 4;

__global__ void coordsum_step_kernel(MagpoCoordSumCfg c, int B, const int32_t* __restrict__ action,
                                     MagpoCoordSumState st, MagpoTimeStep ts, uint8_t* __restrict__ done_out) {
  extern __shared__ int32_t smem_hist[];  // [kStepWarps][time_limit]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kStepWarps + w;
  if (b >= B) return;
  const int A = c.num_agents, a = c.num_actions, TL = c.time_limit;
  int32_t* hist = smem_hist + w * TL;
  for (int i = lane; i < TL; i += 32) hist[i] = 0;
  const int sc = st.step_count[b];
  const int32_t* target = st.target + (size_t)b * (TL + 1);
  const int32_t target_t = target[min(max(sc, 0), TL)];
  int32_t asum = 0;
  for (int i = lane; i < A; i += 32) asum += action[(size_t)b * A + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
  const int32_t a0 = action[(size_t)b * A];
  const bool sum_match = asum == target_t;
  const int row_idx = min(max(target_t, 0), a - 1);
  int32_t* row = st.record + ((size_t)b * a + row_idx) * TL;
  __syncwarp();
  for (int i = lane; i < TL; i += 32) {
    int32_t v = row[i];
    if (v != -1 && v >= 0 && v < TL) atomicAdd(&hist[v], 1);
  }
  __syncwarp();
  // first-max argmax over the TL bins
  int best_c = -1, best_i = 0x7fffffff;
  for (int i = lane; i < TL; i += 32) {
    int cnt = hist[i];
    if (cnt > best_c) { best_c = cnt; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (oc > best_c || (oc == best_c && oi < best_i)) { best_c = oc; best_i = oi; }
  }
  const bool hit = best_i == a0;
  const float reward = sum_match ? (hit ? 1.0f : 2.0f) : 0.0f;
  if (lane == 0) row[min(max(sc, 0), TL - 1)] = a0;
  const int steps = sc + 1;
  const bool done = steps >= TL;
  const int32_t next_target = target[min(steps, TL)];
  // real_next_obs (pre-reset observation)
  coordsum_write_obs(c, b, lane, next_target, steps, ts.next_agents_view, nullptr, ts.next_step_count);
  __syncwarp();
  int32_t obs_target = next_target, obs_step = steps;
  if (done) {
    uint32_t r0, r1;
    prng_split_i(st.key[2 * b], st.key[2 * b + 1], 0u, r0, r1);  // key, _ = split(state.key)
    __syncwarp();
    coordsum_base_reset(c, st, b, lane, r0, r1);
    __syncwarp();
    obs_target = st.target[(size_t)b * (TL + 1)];
    obs_step = 0;
  } else if (lane == 0) {
    st.step_count[b] = steps;
  }
  coordsum_write_obs(c, b, lane, obs_target, obs_step, ts.agents_view, ts.action_mask, ts.step_count);
  for (int i = lane; i < A; i += 32) {
    if (ts.reward) ts.reward[(size_t)b * A + i] = reward;
    if (ts.discount) ts.discount[(size_t)b * A + i] = done ? 0.0f : 1.0f;
  }
  if (lane == 0) {
    // RecordEpisodeMetrics.step; mean(reward) over identical entries == reward (exact in fp32)
    float new_ret = st.running_return[b] + reward;
    int new_len = st.running_length[b] + 1;
    float ep_ret = done ? new_ret : st.episode_return[b];
    int ep_len = done ? new_len : st.episode_length[b];
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

int coordsum_step_launch(cudaStream_t s, const MagpoCoordSumCfg* cfg, int B, const int32_t* action,
                         MagpoCoordSumState st, MagpoTimeStep ts, uint8_t* done_out) {
  if (cfg->num_agents < 1 || cfg->num_actions < 1 || cfg->time_limit < 1 || cfg->maxval < 1) return MAGPO_ERR_ARG;
  if (cfg->num_actions > cfg->time_limit) return MAGPO_ERR_UNSUPPORTED;  // bincount(length=time_limit) would drop entries
  size_t smem = (size_t)kStepWarps * cfg->time_limit * sizeof(int32_t);
  // algorithmic bytes per env-step: record row + target + actions in; obs, mask, reward, metrics out
  const int A_ = cfg->num_agents, a_ = cfg->num_actions;
  ProfScope ps(PROF_ENV, s, (double)B * (4.0 * cfg->time_limit + 12 + 4 * A_ + 8 + 2 * 4 * A_ * (A_ + 1) + a_ * A_ + 16 * A_ + 10));
  coordsum_step_kernel<<<(unsigned)ceil_div(B, kStepWarps), kStepWarps * 32, smem, s>>>(*cfg, B, action, st, ts, done_out);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo

using namespace magpo;

extern "C" {

int magpo_coordsum_reset(magpo_stream_t s, const MagpoCoordSumCfg* cfg, int32_t B, const uint32_t* keys,
                         MagpoCoordSumState st, MagpoTimeStep ts) {
  if (!cfg || !keys || B < 0) return MAGPO_ERR_ARG;
  if (B == 0) return MAGPO_OK;
  coordsum_reset_kernel<<<(unsigned)ceil_div((int64_t)B * 32, 128), 128, 0, as_stream(s)>>>(*cfg, B, keys, st, ts);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_coordsum_step(magpo_stream_t s, const MagpoCoordSumCfg* cfg, int32_t B, const int32_t* action,
                        MagpoCoordSumState st, MagpoTimeStep ts) {
  if (!cfg || !action || B < 0) return MAGPO_ERR_ARG;
  if (B == 0) return MAGPO_OK;
  return coordsum_step_launch(as_stream(s), cfg, B, action, st, ts, nullptr);
}

}  // extern "C"
