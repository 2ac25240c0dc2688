// MAGPO losses, fused forward + gradient w.r.t. both networks' logits and the value.
// Reference: rec_magpo.py:251-311 (_guider_loss_fn: KL(guider || sg(learner)) masked by the double-clip
// region, PPO-clip with the double-clipped ratio, clipped value loss, entropy bonus) and :338-370
// (_actor_loss_fn: KL(sg(guider) || learner) + alpha * PPO-clip); distrax/tfp Categorical arithmetic per
// SURVEY.md Appendix A10; gradients per Appendix G.  One thread per token; a <= 32 actions.
// HBM-bound: per token reads 2*a*4 + a + 24 B, writes 2*a*4 + 4 B.
#include "common.cuh"
#include "kernels.cuh"
#include "update.cuh"

namespace magpo {
namespace {

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in thread 0
}

// Sum (pass 0) or sum of squared deviations (pass 1) of the advantages of each slot's minibatch.
__global__ void __launch_bounds__(256)
adv_stats_kernel(int T, int B, int A, const float* __restrict__ adv, const int32_t* __restrict__ env_index, int n_env,
                 int n_per_slot, int pass, double* __restrict__ acc /*[U][2]*/) {
  __shared__ float sm[8];
  const int slot = blockIdx.y;
  const int64_t per_env = (int64_t)T * A;
  const int64_t total = per_env * n_per_slot;
  const double mean = pass ? acc[slot * 2] / (double)total : 0.0;
  float s = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ne = (int)(idx % n_per_slot);
    const int64_t ta = idx / n_per_slot;
    const int t = (int)(ta / A), i = (int)(ta % A);
    const int li = slot * n_per_slot + ne;
    if (li >= n_env) continue;
    const int b = env_index[li];
    const float v = adv[((int64_t)t * B + b) * A + i];
    if (pass) {
      const float dlt = v - (float)mean;
      s += dlt * dlt;
    } else {
      s += v;
    }
  }
  const float r = block_sum_256(s, sm);
  if (threadIdx.x == 0) atomicAdd(&acc[slot * 2 + pass], (double)r);
}

__global__ void adv_stats_finish_kernel(int U, int64_t total, const double* __restrict__ acc, float* __restrict__ stats) {
  const int u = threadIdx.x;
  if (u >= U) return;
  const double mean = acc[u * 2] / (double)total;
  const double var = acc[u * 2 + 1] / (double)total;
  stats[u * 2] = (float)mean;
  stats[u * 2 + 1] = (float)sqrt(var);  // jnp.std, ddof=0
}

struct LossHyper {
  float clip_eps, ent_coef, vf_coef, lo, hi, alpha, inv_tokens;
};

__global__ void __launch_bounds__(256)
magpo_loss_kernel(int64_t R, int A, int a, LossHyper hp, const float* __restrict__ lg_raw,
                  const float* __restrict__ ll_raw, const uint8_t* __restrict__ mask, const int32_t* __restrict__ action,
                  const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ value,
                  const float* __restrict__ value_old, const float* __restrict__ targets,
                  const int32_t* __restrict__ env_slot, int N, const float* __restrict__ stats,
                  float* __restrict__ dlg, float* __restrict__ dll, float* __restrict__ dvalue,
                  float* __restrict__ loss_sums /*[8]*/) {
  __shared__ float sm[8];
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s_gl = 0.f, s_ent = 0.f, s_vl = 0.f, s_klg = 0.f, s_al = 0.f, s_kla = 0.f;
  if (row < R) {
    const float* g = lg_raw + row * a;
    const float* l = ll_raw + row * a;
    const uint8_t* m = mask + row * a;
    float mg = kF32Min, ml = kF32Min;
    for (int j = 0; j < a; ++j) {
      if (m[j]) { mg = fmaxf(mg, g[j]); ml = fmaxf(ml, l[j]); }
    }
    float sg = 0.f, sl = 0.f;
    for (int j = 0; j < a; ++j) {
      if (m[j]) { sg += expf(g[j] - mg); sl += expf(l[j] - ml); }
    }
    const float lse_g = mg + logf(sg), lse_l = ml + logf(sl);
    float kl_g = 0.f, kl_a = 0.f, ent = 0.f;
    for (int j = 0; j < a; ++j) {
      if (!m[j]) continue;
      const float lpg = g[j] - lse_g, lpl = l[j] - lse_l;
      const float pg = expf(lpg);
      if (pg != 0.f) {
        kl_g += pg * (lpg - lpl);
        ent -= pg * lpg;
      }
    }
    kl_a = kl_g;  // KL(sg(guider) || learner) has the same value; only the gradient differs
    s_kla = kl_a;
    const int act = action[row];
    const bool legal = m[act] != 0;
    const float lgp = (legal ? g[act] : kF32Min) - lse_g;
    const float llp = (legal ? l[act] : kF32Min) - lse_l;
    const float lold = logp_old[row];
    const int n = (int)((row / A) % N);
    const int slot = env_slot[n];
    const float advn = (adv[row] - stats[slot * 2]) / (stats[slot * 2 + 1] + 1e-8f);
    // ---- guider (rec_magpo.py:261-311)
    const float diff = lgp - llp;
    const float ratio = expf(lgp - lold);
    const float cdiff = fminf(fmaxf(diff, hp.lo), hp.hi);
    const float cr = expf(cdiff + llp - lold);
    const float kmask = (diff < hp.lo || diff > hp.hi) ? 1.0f : 0.0f;
    const float lo_r = 1.0f - hp.clip_eps, hi_r = 1.0f + hp.clip_eps;
    const float l1 = ratio * advn;
    const float l2 = fminf(fmaxf(cr, lo_r), hi_r) * advn;
    const float d1 = ratio * advn;
    const float d2 = (cr > lo_r && cr < hi_r && diff > hp.lo && diff < hp.hi) ? cr * advn : 0.0f;
    const float cg = l1 < l2 ? -d1 : (l2 < l1 ? -d2 : -0.5f * (d1 + d2));
    s_gl = -fminf(l1, l2);
    s_klg = kl_g * kmask;
    s_ent = ent;
    // ---- value (rec_magpo.py:297-303)
    const float v = value[row], vo = value_old[row], tg = targets[row];
    const float dv = v - vo;
    const float vclip = vo + fminf(fmaxf(dv, -hp.clip_eps), hp.clip_eps);
    const float e1 = v - tg, e2 = vclip - tg;
    const float a1 = e1 * e1, a2 = e2 * e2;
    const float in2 = (dv > -hp.clip_eps && dv < hp.clip_eps) ? e2 : 0.0f;
    const float gv = a1 > a2 ? e1 : (a2 > a1 ? in2 : 0.5f * (e1 + in2));
    s_vl = 0.5f * fmaxf(a1, a2);
    dvalue[row] = hp.inv_tokens * hp.vf_coef * gv;
    // ---- learner (rec_magpo.py:340-370)
    const float ra = expf(llp - lold);
    const float b1 = ra * advn;
    const float b2 = fminf(fmaxf(ra, lo_r), hi_r) * advn;
    const float e2a = (ra > lo_r && ra < hi_r) ? ra * advn : 0.0f;
    const float ca = b1 < b2 ? -b1 : (b2 < b1 ? -e2a : -0.5f * (b1 + e2a));
    s_al = -fminf(b1, b2);
    // ---- gradients w.r.t. the (masked) logits
    for (int j = 0; j < a; ++j) {
      float og = 0.f, ol = 0.f;
      if (m[j]) {
        const float lpg = g[j] - lse_g, lpl = l[j] - lse_l;
        const float pg = expf(lpg), pl = expf(lpl);
        const float onehot = (j == act) ? 1.0f : 0.0f;
        og = cg * (onehot - pg) + kmask * pg * ((lpg - lpl) - kl_g) + hp.ent_coef * pg * (lpg + ent);
        ol = (pl - pg) + hp.alpha * ca * (onehot - pl);
      }
      dlg[row * a + j] = hp.inv_tokens * og;
      dll[row * a + j] = hp.inv_tokens * ol;
    }
  }
  const float r1 = block_sum_256(s_gl, sm);
  const float r2 = block_sum_256(s_ent, sm);
  const float r3 = block_sum_256(s_vl, sm);
  const float r4 = block_sum_256(s_klg, sm);
  const float r6 = block_sum_256(s_al, sm);
  const float r7 = block_sum_256(s_kla, sm);
  if (threadIdx.x == 0) {
    atomicAdd(loss_sums + 1, r1 * hp.inv_tokens);
    atomicAdd(loss_sums + 2, r2 * hp.inv_tokens);
    atomicAdd(loss_sums + 3, r3 * hp.inv_tokens);
    atomicAdd(loss_sums + 4, r4 * hp.inv_tokens);
    atomicAdd(loss_sums + 6, r6 * hp.inv_tokens);
    atomicAdd(loss_sums + 7, r7 * hp.inv_tokens);
  }
}

}  // namespace

int adv_stats(cudaStream_t s, int T, int B, int A, const float* adv, const int32_t* env_index, int n_env, int U,
              double* acc /*[U][2] scratch*/, float* stats /*[U][2]*/) {
  if (U < 1 || U > 32 || n_env % U) return MAGPO_ERR_ARG;
  const int nps = n_env / U;
  const int64_t total = (int64_t)T * A * nps;
  ProfScope ps(PROF_LOSS, s, 8.0 * (double)total * U);
  MAGPO_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * U, s));
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(total, 256), 2 * kNumSMs), U);
  adv_stats_kernel<<<grid, 256, 0, s>>>(T, B, A, adv, env_index, n_env, nps, 0, acc);
  MAGPO_LAUNCH_OK();
  adv_stats_kernel<<<grid, 256, 0, s>>>(T, B, A, adv, env_index, n_env, nps, 1, acc);
  MAGPO_LAUNCH_OK();
  adv_stats_finish_kernel<<<1, 32, 0, s>>>(U, total, acc, stats);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

int magpo_losses(cudaStream_t s, int64_t R, int N, int A, int a, const MagpoSysCfg* sys, float inv_tokens,
                 const float* lg, const float* ll, const uint8_t* mask, const int32_t* action, const float* logp_old,
                 const float* adv, const float* value, const float* value_old, const float* targets,
                 const int32_t* env_slot, const float* stats, float* dlg, float* dll, float* dvalue,
                 float* loss_sums) {
  if (R <= 0) return MAGPO_OK;
  ProfScope ps(PROF_LOSS, s, (double)R * (16.0 * a + a + 32.0));
  LossHyper hp;
  hp.clip_eps = (float)sys->clip_eps;
  hp.ent_coef = (float)sys->ent_coef;
  hp.vf_coef = (float)sys->vf_coef;
  hp.lo = (float)log(1.0 / sys->clip_gpo);
  hp.hi = (float)log(sys->clip_gpo);
  hp.alpha = (float)sys->alpha;
  hp.inv_tokens = inv_tokens;
  magpo_loss_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, s>>>(R, A, a, hp, lg, ll, mask, action, logp_old, adv, value,
                                                               value_old, targets, env_slot, N, stats, dlg, dll, dvalue,
                                                               loss_sums);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

}  // namespace magpo
