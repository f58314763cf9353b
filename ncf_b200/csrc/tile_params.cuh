// Kernel-argument block shared by the forward / training-step kernels.
#pragma once
#include "common.cuh"

struct TileParams {
  // model (device pointers)
  int type, f, d, L, predict_size;
  int W[NCF_MAX_LAYERS + 1];
  const float *eug, *eig, *eum, *eim;
  const float* w[NCF_MAX_LAYERS];
  const float* b[NCF_MAX_LAYERS];
  const float *pw, *pb;
  int64_t U, I;
  // batch
  const int64_t* user;
  const int64_t* item;
  int user_div;          // >0: sample b uses user[b / user_div] (evaluation: one user, C candidates)
  const float* label;    // NULL => inference
  const float* teacher;  // NULL => plain BCE
  const float* dlogit_in; // non-NULL => use this dloss/dlogit instead of evaluating the loss
  float alpha, invB;
  int64_t B;
  float* logits;         // nullable
  double* loss_accum;
  // gradients
  float *gug, *gig, *gum, *gim, *gt;
  int32_t *uflag, *iflag, *tcount;
  int64_t *ulist, *ilist;
  int64_t w_off[NCF_MAX_LAYERS], b_off[NCF_MAX_LAYERS], pw_off, pb_off;
  // HBM scratch of per-sample activations / deltas (generic path)
  float* act[NCF_MAX_LAYERS + 1];    // act[k],   k = 1..L-1 : [B, W[k]]
  float* delta[NCF_MAX_LAYERS + 1];  // delta[k], k = 1..L   : [B, W[k]]
  // pre-split (hi, lo) TF32 weights for the mma path: forward operand and transposed operand
  float* wsplit_f[NCF_MAX_LAYERS];
  float* wsplit_b[NCF_MAX_LAYERS];
  // shared-memory layout (float offsets), filled by the launcher
  int smem_off[NCF_MAX_LAYERS + 1];
  int gmf_off, stage_off, misc_off;
};

namespace ncf {

inline void fill_model_params(TileParams& p, const NcfModel* m) {
  TowerShape ts = make_tower_shape(m->model_type, m->factor_num, m->num_layers);
  p.type = m->model_type;
  p.f = m->factor_num;
  p.d = m->mlp_dim;
  p.L = m->num_layers;
  p.predict_size = ts.predict_size;
  for (int k = 0; k <= p.L; ++k) p.W[k] = ts.width[k];
  p.eug = m->embed_user_gmf;
  p.eig = m->embed_item_gmf;
  p.eum = m->embed_user_mlp;
  p.eim = m->embed_item_mlp;
  for (int k = 0; k < p.L; ++k) {
    p.w[k] = m->mlp_w[k];
    p.b[k] = m->mlp_b[k];
    p.w_off[k] = ts.w_off[k];
    p.b_off[k] = ts.b_off[k];
  }
  p.pw = m->predict_w;
  p.pb = m->predict_b;
  p.pw_off = ts.pw_off;
  p.pb_off = ts.pb_off;
  p.U = m->user_num;
  p.I = m->item_num;
}

inline void fill_grad_params(TileParams& p, const NcfGrads* g) {
  p.gug = g->g_user_gmf;
  p.gig = g->g_item_gmf;
  p.gum = g->g_user_mlp;
  p.gim = g->g_item_mlp;
  p.gt = g->g_tower;
  p.uflag = g->user_flag;
  p.iflag = g->item_flag;
  p.ulist = g->user_list;
  p.ilist = g->item_list;
  p.tcount = g->touched_count;
}

int launch_generic_forward(TileParams& p, cudaStream_t st);
int launch_generic_train(TileParams& p, cudaStream_t st);

// tensor-pipe path (tile_mma.cu): eligible when mma_tile_rows(p) != 0
int mma_tile_rows(const TileParams& p);
int64_t mma_split_floats(const TileParams& p);
int mma_prepare_weights(TileParams& p, float* ws, cudaStream_t st);
int launch_mma_forward(TileParams& p, int passes, cudaStream_t st);
int launch_mma_train(TileParams& p, int passes, cudaStream_t st);
inline int tower_passes(const NcfModel* m) { return m->tower_math == NCF_MATH_TF32 ? 1 : 3; }

// narrow towers, one thread per sample on the FMA pipe (tile_small.cu): factor_num 8, 1..3 layers
bool small_eligible(const TileParams& p);
int launch_small_forward(TileParams& p, cudaStream_t st);
int launch_small_train(TileParams& p, cudaStream_t st);

// wide towers at small batches, forward only: one launch per layer over many CTAs (tile_wide.cu)
bool wide_eligible(const TileParams& p);  // reads p.B
int64_t wide_workspace_floats(const TileParams& p, int64_t B);
int launch_wide_forward(TileParams& p, float* ws, cudaStream_t st);

// tcgen05 path (tile_umma.cu): large batches, tower widths that are power-of-two multiples of 32
bool umma_eligible(const TileParams& p);  // reads p.B
int64_t umma_forward_workspace_floats(const TileParams& p, int64_t B);
int64_t umma_train_workspace_floats(const TileParams& p, int64_t B);
int launch_umma_forward(TileParams& p, int passes, float* ws, cudaStream_t st);
int launch_umma_train(TileParams& p, int passes, float* ws, cudaStream_t st);

}  // namespace ncf
