// General Sable shapes: network.net_config.{embed_dim, n_head, n_block} beyond the default (64, 1, 1) — the reference loops over blocks
// and heads (sable_network.py:111-119,286-294, retention.py:229-259,281-287) and its tuned runs use embed_dim 32 / 64 / 128, n_head 1 / 2 /
// 4, n_block 1 / 2 / 3 (experiment_data/params.csv:61-103). The default shape keeps its specialised kernels (sable.cu, rowops.cu,
// retention_chunk.cu, sable_step.cu); every other shape runs the layer-by-layer path declared here: the same GEMMs (tcgen05 3xTF32 where the
// shape fits, fp32 SIMT otherwise), row kernels templated on the row width, and a per-(env, head) retention scan.
#pragma once
#include <vector>

#include "params.cuh"

namespace magpo {

struct HeadKappas {  // per-head decay factors, passed to the scan kernels by value
  float k[4];
};

struct NetShape {
  int D, nh, nb, hs;  // embed_dim, n_head, n_block, head size D / nh
  int A, d, a, max_step;
  static NetShape of(const MagpoNetCfg* n) {
    NetShape s;
    s.D = n->embed_dim; s.nh = n->n_head; s.nb = n->n_block; s.hs = n->embed_dim / n->n_head;
    s.A = n->n_agents; s.d = n->obs_dim; s.a = n->action_dim; s.max_step = n->max_step_count;
    return s;
  }
  bool is_default() const { return D == kD && nh == 1 && nb == 1; }
  int64_t state_elems() const { return (int64_t)nh * nb * hs * hs; }  // one Sable state of one env: [n_head, n_block, hs, hs]
};
// shapes the general path covers
inline bool net_shape_ok(const MagpoNetCfg* n) {
  const int D = n->embed_dim, nh = n->n_head, nb = n->n_block;
  if (D != 32 && D != 64 && D != 128) return false;
  if (nh != 1 && nh != 2 && nh != 4) return false;
  if (nb < 1 || nb > 3) return false;
  return D % nh == 0 && (D / nh) % nh == 0;  // GroupNorm(num_groups = n_head) over the head size (retention.py:247,289)
}

// decay kappa of head h (retention.py:231-234): (1 - exp(linspace(log(1/32), log(1/512), n_head)[h])) * decay_scaling_factor, float32
inline float head_kappa(const MagpoNetCfg* n, int h) {
  // jnp.linspace in float32: start + i * ((stop - start) / (num - 1)), the last point set to stop
  const float lo = logf(1.0f / 32.0f), hi = logf(1.0f / 512.0f);
  float x = lo;
  if (n->n_head > 1) x = (h == n->n_head - 1) ? hi : lo + (float)h * ((hi - lo) / (float)(n->n_head - 1));
  return (1.0f - expf(x)) * n->decay_scaling_factor;
}

// Flat parameter layout of the general guider: the default layout of params.cuh with the per-block groups repeated n_block times
// (for (64, 1, 1) the two layouts coincide). Packed groups as there: [w_q(heads) | w_k(heads) | w_v(heads) | w_g] is one [D, 4D]
// matrix — head h's w_q is columns [h hs, (h+1) hs) of the first D columns — and [W_gate | W_linear] one [D, 2D] matrix.
struct RetnG {
  float *qkvg, *wo, *gn_s, *gn_b;
};
struct EncBlockG {
  float *ln1, *ln2;
  RetnG r;
  float *ffn_gl, *ffn_out;
};
struct DecBlockG {
  float *ln1, *ln2, *ln3;
  RetnG r1, r2;
  float *ffn_gl, *ffn_out;
};
struct GuiderG {
  float *obs_scale, *Wobs, *ln;
  EncBlockG enc[3];
  float *h0_w, *h0_b, *h2_s, *h3_w, *h3_b;
  float *Wa, *dln;
  DecBlockG dec[3];
  float *dh0_w, *dh0_b, *dh2_s, *dh3_w, *dh3_b;
  int64_t total;
  static GuiderG bind(float* base, const NetShape& s);
};

struct ParamEntryG {
  char name[96];
  int64_t offset;
  int32_t dim0, dim1, ld;
};
// flax tree paths of the general guider (retention_heads_{h}/w_q ... as strided views of the packed matrices)
void guider_table_g(const NetShape& s, std::vector<ParamEntryG>* out);

// One forward / backward of the general guider over R = T*N*A token rows (time-major). Workspace planned by the caller.
struct SableActsG;
struct GuiderGT;
size_t sable_g_workspace_bytes(const NetShape& s, int T, int N, bool with_backward);
// training forward (SableNetwork.__call__): value [R], raw logits [R, a]; states h_* are [N, nh, nb, hs, hs]
int sable_g_train_forward(cudaStream_t st, const MagpoNetCfg* net, const float* guider, int T, int N, const float* agents_view,
                          const int32_t* step_count, const uint8_t* done, const int32_t* action, const float* h_enc, const float* h_self,
                          const float* h_cross, float* value, float* logits, void* ws, size_t ws_bytes, bool with_backward);
// backward of the forward that last ran in `ws` (with_backward = true); grads accumulate into the flat buffer g
int sable_g_train_backward(cudaStream_t st, const MagpoNetCfg* net, const float* guider, int T, int N, const float* agents_view,
                           const int32_t* step_count, const uint8_t* done, const int32_t* action, const float* h_enc, const float* h_self,
                           const float* h_cross, const float* dlogits, const float* dvalue, float* g, void* ws, size_t ws_bytes);
// SableNetwork.get_actions for B envs, one timestep (layer-by-layer; states updated in place unless dry)
size_t sable_g_step_workspace_bytes(const NetShape& s, int B);
int sable_g_get_actions(cudaStream_t st, const MagpoNetCfg* net, int B, int gumbel_rows, const float* guider, const float* agents_view,
                        const uint8_t* action_mask, const int32_t* step_count, const uint8_t* prev_done, const uint32_t* sample_keys,
                        MagpoSableHState hs, bool dry, int32_t* action, float* log_prob, float* value, float* masked_logits, void* ws,
                        size_t ws_bytes, bool prepare);
// PE table + transposed weights + TF32 images of the step workspace (what `prepare = true` does inside sable_g_get_actions)
int sable_g_prepare(cudaStream_t st, const MagpoNetCfg* net, int B, const float* guider, void* ws, size_t ws_bytes);

// ---- generic_rows.cu: row kernels on D-wide rows (D = 32 VPL, VPL in {1, 2, 4}); same contracts as their 64-wide namesakes in kernels.cuh
int g_act_rms_fwd(cudaStream_t s, int D, int64_t R, const float* z, const float* res, const float* scale, int flags, const float* pe,
                  const int32_t* step, int max_step, float* y, float* ype);
int g_act_rms_bwd(cudaStream_t s, int D, int64_t R, const float* z, const float* res, const float* scale, int flags, const float* dy1,
                  const float* dy2, const float* dy3, float* dout, float* dscale);
// gated = swish(g) * GroupNorm_{n_head groups per head row}(ret): ret rows are [n_head, hs]; scale / bias [hs] (retention.py:289-295)
int g_gn_gate_fwd(cudaStream_t s, int D, int nh, int64_t R, const float* g, int ldg, const float* ret, const float* gs, const float* gb,
                  float* gated);
int g_gn_gate_bwd(cudaStream_t s, int D, int nh, int64_t R, const float* g, int ldg, const float* ret, const float* gs, const float* gb,
                  const float* dgated, float* dg, int lddg, float* dret, float* dgs, float* dgb);
int g_swiglu_fwd(cudaStream_t s, int D, int64_t R, const float* gl, float* h);
int g_swiglu_bwd(cudaStream_t s, int D, int64_t R, const float* gl, const float* dh, float* dgl);
int g_embed_fwd(cudaStream_t s, int D, int64_t R, int A, const int32_t* action, const float* Wa, const float* scale, const float* pe,
                const int32_t* step, int max_step, float* x, float* xpe);
int g_embed_bwd(cudaStream_t s, int D, int64_t R, int A, const int32_t* action, const float* Wa, const float* scale, const float* dy1,
                const float* dy2, float* dWa, float* dscale);
int g_pe_table(cudaStream_t s, int D, int max_step, float* pe, bool enabled);
// y[r] += b (bias add over rows), y = a + b, and y[r, :] = gelu / plain pass-through helpers used by the heads
int g_add_rows(cudaStream_t s, int64_t n, const float* a, const float* b, float* y);  // y = a + b (elementwise, n floats)
// per-(env, head) retention scan over T steps; q, k, v: column blocks of a packed buffer (row stride ld), head h at columns [h hs, (h+1) hs)
// H0 / Hout [N, nh, nb, hs, hs] (block `blk`), Hsave [T, N, nh, hs, hs] (state after each timestep, backward only)
int g_retention_fwd(cudaStream_t s, const NetShape& sh, const HeadKappas& kappas, int blk, int T, int N, int rows_per_step, bool causal,
                    const float* q, const float* k, const float* v, int ld, const float* H0, const uint8_t* done, float* ret, int ldr,
                    float* Hsave, float* Hout);
int g_retention_bwd(cudaStream_t s, const NetShape& sh, const HeadKappas& kappas, int T, int N, int rows_per_step, bool causal, const float* q,
                    const float* k, const float* v, int ld, const uint8_t* done, const float* Hsave, const float* dret, int ldr, float* dq,
                    float* dk, float* dv, int ldd);

}  // namespace magpo
