// Sable guider (SableNetwork, networks/sable_network.py:346-482): training forward/backward over a
// time-major token batch and the per-timestep inference path, as sequences of the kernels in kernels.cuh.
#pragma once
#include "params.cuh"

namespace magpo {

// Transposed copies of the weight matrices, refreshed once per optimiser step (used for dX = dY @ W^T).
struct GuiderT {
  float *WobsT, *qkvgT, *woT, *ffn_glT, *ffn_outT, *h0T;
  float *qkvg1T, *wo1T, *qkvg2T, *wo2T, *dffn_glT, *dffn_outT, *dh0T;
  float *region_hi, *region_lo;  // TF32 hi/lo split of [WobsT, dh0T + 64*64)
  int64_t region_n;
  void plan(Arena& ar, int d);
};
int guider_transpose(cudaStream_t s, const GuiderP& p, const GuiderT& t, int d);

// Saved activations of one training forward over R = T*N*A rows (all fp32, row-major).
struct SableActs {
  // encoder
  float *on, *z0, *xin, *kqv, *qkvg, *ret, *gated, *o, *x1, *gl, *hmid, *f, *x, *xpe, *zh;
  // decoder
  float *xD, *xpeD, *qkvg1, *ret1, *gated1, *o1, *rpe, *qkvg2, *ret2, *gated2, *o2, *y, *glD, *hmidD, *fD, *xd, *zhD;
  // per-timestep retention states [T,N,64,64] (backward only)
  float *Hs_enc, *Hs_self, *Hs_cross;
  // scratch for the backward
  float *tA, *tB, *tC, *tD, *tE, *tQ, *tG, *t_d;
  void plan(Arena& ar, int64_t R, int64_t TN, int d, bool with_backward);
};

struct SableBatch {
  int T, N, A, d, a, max_step;
  const float* agents_view;   // [T,N,A,d]
  const int32_t* step_count;  // [T,N,A]
  const uint8_t* done;        // [T,N]
  const int32_t* action;      // [T,N,A]
  const float *h_enc, *h_self, *h_cross;  // [N,64,64] states at the start of the sequence
  const float* pe;            // [max_step+1, 64]
  float kappa;
};

// SableNetwork.__call__ (sable_network.py:412-441): value [R], raw (un-masked) logits [R,a].
int sable_train_forward(cudaStream_t s, const GuiderP& p, const GuiderT* pt, const SableBatch& b, const SableActs& w,
                        float* value, float* logits, bool save_states);
// Backward of the above given dL/dlogits [R,a] (zero at illegal actions) and dL/dvalue [R]; grads accumulate into g.
int sable_train_backward(cudaStream_t s, const GuiderP& p, const GuiderT& pt, const SableBatch& b, const SableActs& w,
                         const float* dlogits, const float* dvalue, const GuiderP& g);

}  // namespace magpo
