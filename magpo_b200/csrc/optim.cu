// optax.chain(clip_by_global_norm(max_norm), adam(lr, eps=1e-5)) + optax.apply_updates, fused.
// Reference: rec_magpo.py:581-589,412-423; optax 0.2.4 arithmetic per SURVEY.md Appendix A11.
// Two passes over the flat buffers: (1) sum of squares -> global norm, (2) moments + parameter update.
// HBM-bound: pass 1 reads 4 B/param, pass 2 reads 16 B and writes 12 B per parameter.
#include "common.cuh"
#include "kernels.cuh"

namespace magpo {
namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(int64_t n, const float* __restrict__ g, float scale, float* __restrict__ out) {
  __shared__ float sm[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i] * scale;
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) atomicAdd(out, r);
  }
}

// scratch[0] = sum of squares (in) ; out: scratch[1] = clip factor numerator flag, [2] = norm, [3] = bc1, [4] = bc2, [5] = lr
// Learning rate (utils/training.py:30-64, make_learning_rate): constant, or with `decay_learning_rates` the linear schedule
// init_lr * (1 - (count // (ppo_epochs * num_minibatches)) / num_updates) evaluated at the optimiser count BEFORE the increment
// (optax scale_by_schedule); the count lives on the device, so the schedule needs no host synchronisation.
__global__ void adam_prelude_kernel(float* __restrict__ scratch, int32_t* __restrict__ count, float max_norm, float init_lr,
                                    int32_t decay_period, int32_t num_updates) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float norm = sqrtf(scratch[0]);
  int32_t c = *count;
  float lr = init_lr;
  if (decay_period > 0) {
    const float frac = 1.0f - (float)(c / decay_period) / (float)num_updates;
    lr = init_lr * frac;
  }
  scratch[5] = lr;
  c = c < 2147483647 ? c + 1 : c;  // optax safe_int32_increment
  *count = c;
  scratch[1] = (norm < max_norm) ? 0.0f : 1.0f;  // 1 => rescale by (g / norm) * max_norm
  scratch[2] = norm;
  scratch[3] = 1.0f - powf(0.9f, (float)c);
  scratch[4] = 1.0f - powf(0.999f, (float)c);
}

__global__ void __launch_bounds__(256)
adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
            float* __restrict__ nu, const float* __restrict__ scratch, float grad_scale, float max_norm) {
  const float clip = scratch[1], norm = scratch[2], bc1 = scratch[3], bc2 = scratch[4], lr = scratch[5];
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-5f;
  const float omb1 = (float)(1.0 - 0.9), omb2 = (float)(1.0 - 0.999);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (clip != 0.0f) gi = (gi / norm) * max_norm;
    const float m = omb1 * gi + b1 * mu[i];
    const float v = omb2 * (gi * gi) + b2 * nu[i];
    mu[i] = m;
    nu[i] = v;
    const float u = (m / bc1) / (sqrtf(v / bc2) + eps);
    p[i] = p[i] + (-lr) * u;
  }
}

}  // namespace
}  // namespace magpo

using namespace magpo;

extern "C" int magpo_clip_adam_sched(magpo_stream_t s_, int64_t n, float* params, const float* grads, float* mu, float* nu,
                                     int32_t* count, float grad_scale, float lr, int32_t decay_period, int32_t num_updates,
                                     float max_norm, float* scratch) {
  if (n < 0 || !params || !grads || !mu || !nu || !count || !scratch) return MAGPO_ERR_ARG;
  if (decay_period < 0 || (decay_period > 0 && num_updates < 1)) return MAGPO_ERR_ARG;
  if (n == 0) return MAGPO_OK;
  cudaStream_t s = as_stream(s_);
  ProfScope ps(PROF_OPTIM, s, 32.0 * (double)n);
  MAGPO_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 8, s));
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 4);
  sumsq_kernel<<<grid, 256, 0, s>>>(n, grads, grad_scale, scratch);
  MAGPO_LAUNCH_OK();
  adam_prelude_kernel<<<1, 32, 0, s>>>(scratch, count, max_norm, lr, decay_period, num_updates);
  MAGPO_LAUNCH_OK();
  adam_kernel<<<grid, 256, 0, s>>>(n, params, grads, mu, nu, scratch, grad_scale, max_norm);
  MAGPO_LAUNCH_OK();
  return MAGPO_OK;
}

extern "C" int magpo_clip_adam(magpo_stream_t s_, int64_t n, float* params, const float* grads, float* mu, float* nu,
                               int32_t* count, float grad_scale, float lr, float max_norm, float* scratch) {
  return magpo_clip_adam_sched(s_, n, params, grads, mu, nu, count, grad_scale, lr, 0, 0, max_norm, scratch);
}
