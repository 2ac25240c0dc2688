// Flat parameter layouts of the two networks (shared by params, grads and Adam moments).
// Tensors that always enter the same GEMM are stored packed side by side so that one GEMM serves them:
//   retention  [w_q | w_k | w_v | w_g]  as one [64,256] matrix (retention.py:50-64,237-241)
//   SwiGLU     [W_gate | W_linear]      as one [64,128] matrix (torsos.py:89-93)
//   GRU        [ir | iz | in] kernels [128,384] + biases [384], [hr | hz | hn] kernels [128,384] (flax GRUCell, App. A8)
// magpo_param_tensor() reports every flax tensor as (offset, rows, cols, row stride) into the flat buffer,
// so the host exposes them under their flax tree paths as strided views. Offsets are 16-byte aligned;
// padding elements stay zero (zero gradient -> zero Adam update).
#pragma once
#include "kernels.cuh"

namespace magpo {

// X(field, rows, cols): rows*cols floats, row-major contiguous.
#define MAGPO_GUIDER_FIELDS(X, d, a)                                                              \
  X(obs_scale, 1, d) X(Wobs, d, kD) X(ln, 1, kD) X(ln1, 1, kD) X(ln2, 1, kD) X(qkvg, kD, 4 * kD)  \
  X(wo, kD, kD) X(gn_s, 1, kD) X(gn_b, 1, kD) X(ffn_gl, kD, 2 * kD) X(ffn_out, kD, kD)            \
  X(h0_w, kD, kD) X(h0_b, 1, kD) X(h2_s, 1, kD) X(h3_w, kD, 1) X(h3_b, 1, 1)                      \
  X(Wa, (a) + 1, kD) X(dln, 1, kD) X(dln1, 1, kD) X(dln2, 1, kD) X(dln3, 1, kD)                   \
  X(qkvg1, kD, 4 * kD) X(wo1, kD, kD) X(gn1_s, 1, kD) X(gn1_b, 1, kD)                             \
  X(qkvg2, kD, 4 * kD) X(wo2, kD, kD) X(gn2_s, 1, kD) X(gn2_b, 1, kD)                             \
  X(dffn_gl, kD, 2 * kD) X(dffn_out, kD, kD) X(dh0_w, kD, kD) X(dh0_b, 1, kD) X(dh2_s, 1, kD)     \
  X(dh3_w, kD, a) X(dh3_b, 1, a)

#define MAGPO_ACTOR_FIELDS(X, d, a)                                                               \
  X(pre_w, d, kH) X(pre_b, 1, kH) X(Wi, kH, 3 * kH) X(bi, 1, 3 * kH) X(Wh, kH, 3 * kH)            \
  X(bhn, 1, kH) X(post_w, kH, kH) X(post_b, 1, kH) X(head_w, kH, a) X(head_b, 1, a)

inline int64_t align4(int64_t n) { return (n + 3) & ~int64_t(3); }

struct GuiderP {
#define X(f, r, c) float* f;
  MAGPO_GUIDER_FIELDS(X, 0, 0)
#undef X
  int64_t total;
  static GuiderP bind(float* base, int d, int a) {
    GuiderP p;
    int64_t off = 0;
#define X(f, r, c) p.f = base ? base + off : nullptr; off += align4((int64_t)(r) * (c));
    MAGPO_GUIDER_FIELDS(X, d, a)
#undef X
    p.total = off;
    return p;
  }
};

struct ActorP {
#define X(f, r, c) float* f;
  MAGPO_ACTOR_FIELDS(X, 0, 0)
#undef X
  int64_t total;
  static ActorP bind(float* base, int d, int a) {
    ActorP p;
    int64_t off = 0;
#define X(f, r, c) p.f = base ? base + off : nullptr; off += align4((int64_t)(r) * (c));
    MAGPO_ACTOR_FIELDS(X, d, a)
#undef X
    p.total = off;
    return p;
  }
};

inline int check_net(const MagpoNetCfg* n) {
  if (!n) return MAGPO_ERR_ARG;
  if (n->hidden != kH) return MAGPO_ERR_UNSUPPORTED;
  {  // guider shapes: the default (64, 1, 1) on the specialised kernels, the others on the general path (generic.cuh: net_shape_ok)
    const int D = n->embed_dim, nh = n->n_head, nb = n->n_block;
    if (D != 32 && D != 64 && D != 128) return MAGPO_ERR_UNSUPPORTED;
    if (nh != 1 && nh != 2 && nh != 4) return MAGPO_ERR_UNSUPPORTED;
    if (nb < 1 || nb > 3 || D % nh || (D / nh) % nh) return MAGPO_ERR_UNSUPPORTED;
  }
  if (n->n_agents < 1 || n->n_agents > kMaxAgents) return MAGPO_ERR_UNSUPPORTED;
  if (n->action_dim < 1 || n->action_dim > kMaxActions) return MAGPO_ERR_UNSUPPORTED;
  if (n->obs_dim < 1 || n->obs_dim > 128) return MAGPO_ERR_UNSUPPORTED;
  if (n->max_step_count < 0 || n->max_step_count > 65535) return MAGPO_ERR_ARG;
  return MAGPO_OK;
}

// decay kappa for the single head: (1 - exp(log(1/32))) * decay_scaling_factor in float32
// (retention.py:231-234, sable_network.py:366-369).
inline float net_kappa(const MagpoNetCfg* n) {
  const float e = expf(logf(1.0f / 32.0f));
  return (1.0f - e) * n->decay_scaling_factor;
}

}  // namespace magpo
