// Learner policy (RecurrentActor, networks/base.py:152-184) sequencers.
#pragma once
#include "params.cuh"

namespace magpo {

struct ActorT {
  float *WiT, *WhT, *postT, *headT;
  float *region_hi, *region_lo;  // TF32 hi/lo split of [WiT, headT + a*128)
  int64_t region_n;
  void plan(Arena& ar, int a);
};
int actor_transpose(cudaStream_t s, const ActorP& p, const ActorT& t, int a);

// R = T*Rs rows, Rs = N*A rows per timestep.
struct ActorActs {
  float *e, *gi, *gh, *HU, *Y, *post;          // forward
  float *rzn, *ghn, *dgh, *dA, *dB, *carry;    // saved for / scratch of the backward
  void plan(Arena& ar, int64_t R, int64_t Rs, int a, bool with_backward);
};

int actor_forward(cudaStream_t s, const ActorP& p, const ActorT* pt, int T, int N, int A, int d, int a, const float* agents_view,
                  const uint8_t* done, const float* h0, const ActorActs& w, float* logits, float* h_out);
int actor_backward(cudaStream_t s, const ActorP& p, const ActorT& pt, int T, int N, int A, int d, int a,
                   const float* agents_view, const uint8_t* done, const ActorActs& w, const float* dlogits,
                   const ActorP& g);

}  // namespace magpo
