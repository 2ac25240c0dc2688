// Pieces of the update step (rec_magpo.py:214-499) shared between loss.cu and update.cu.
#pragma once
#include "kernels.cuh"

namespace magpo {

// Per-slot mean / population std of the un-normalised advantages of one minibatch (rec_magpo.py:283,356).
int adv_stats(cudaStream_t s, int T, int B, int A, const float* adv, const int32_t* env_index, int n_env, int U,
              double* acc, float* stats);

int magpo_losses(cudaStream_t s, int64_t R, int N, int A, int a, const MagpoSysCfg* sys, float inv_tokens,
                 const float* lg, const float* ll, const uint8_t* mask, const int32_t* action, const float* logp_old,
                 const float* adv, const float* value, const float* value_old, const float* targets,
                 const int32_t* env_slot, const float* stats, float* dlg, float* dll, float* dvalue, float* loss_sums);

// guider / learner work may run on forked streams (magpo_debug_set_overlap)
bool nets_overlap_enabled();

}  // namespace magpo
