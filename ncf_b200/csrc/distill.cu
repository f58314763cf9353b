// Knowledge-distillation terms beyond the teacher-score MSE that the fused training kernel evaluates in
// its own epilogue (reference src/distillation/base.py:26-50, response.py:34-61, feature.py:48-146).
//
//   ncf_loss_grad_kd   per-sample loss and dloss/dlogit of
//                        w_task * BCE(x, y) + w_kd * KD(x, t)
//                      KD = (x - t)^2                                   (kd_mode 0: ResponseDistillation)
//                      KD = T^2 * (sigmoid(x/T) - sigmoid(t/T))^2       (kd_mode 1: BaseDistillation /
//                           SoftTargetDistillation / the response term of Feature- and AttentionDistillation)
//                      both means over the batch; the result feeds ncf_backward as dlogit.
//   ncf_feature_kd     feature matching on the two embedding-level features of feature.py:56-67:
//                        gmf_features = E_uG[u] * E_iG[i]      mlp_input = [E_uM[u] ; E_iM[i]]
//                      loss += weight * mean_{b, j} (A s_b + a - t_b)_j^2 with the (fixed, never optimised:
//                      scripts/train_student.py:131) adapter Linear (A, a) when student and teacher widths
//                      differ, identity otherwise; its gradient goes straight into the student's
//                      embedding-gradient rows (vector REDs, the rows are already registered by the step).
// One warp per sample; the adapters are at most a few hundred KB and stay L2-resident.
#include "common.cuh"

namespace {

__global__ void loss_grad_kd_kernel(const float* __restrict__ logits, const float* __restrict__ label,
                                    const float* __restrict__ teacher, float w_task, float w_kd, float T, int kd_mode,
                                    int64_t B, float invB, double* __restrict__ loss_accum, float* __restrict__ dlogit) {
  float ls = 0.f;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float x = logits[b], y = label[b];
    const float e = expf(-fabsf(x));
    const float bce = fmaxf(x, 0.f) - x * y + log1pf(e);
    const float sig = (x >= 0.f) ? 1.f / (1.f + e) : e / (1.f + e);
    float l = w_task * bce, d = w_task * (sig - y);
    if (teacher != nullptr && w_kd != 0.f) {
      const float t = teacher[b];
      if (kd_mode == 0) {
        const float df = x - t;
        l += w_kd * df * df;
        d += w_kd * 2.f * df;
      } else {
        const float ss = 1.f / (1.f + expf(-x / T)), st = 1.f / (1.f + expf(-t / T));
        const float df = ss - st;
        l += w_kd * T * T * df * df;
        d += w_kd * 2.f * T * df * ss * (1.f - ss);   // T^2 * 2 df * ss (1 - ss) / T
      }
    }
    ls += l;
    if (dlogit != nullptr) dlogit[b] = d * invB;
  }
  __shared__ float part[32];
  ls = warp_sum(ls);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ls;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss_accum != nullptr) atomicAdd(loss_accum, (double)v * (double)invB);
  }
}

constexpr int kFeatWarps = 4;
constexpr int kMaxFeat = 2048;   // widest feature vector (teacher side) held per warp in shared memory

// s_u, s_i: the student's rows; kind 0: feature = s_u * s_i (width ds), kind 1: feature = [s_u ; s_i] (width 2 ds).
// A: [dt_out, ds_in] row-major adapter (nullptr = identity, needs equal widths), a: [dt_out] bias.
// t_u, t_i: the teacher's rows (feature built the same way, width dt_out).
__global__ void __launch_bounds__(kFeatWarps * 32)
feature_kd_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ item, int64_t B, int64_t U, int64_t I,
                  int kind, const float* __restrict__ su, const float* __restrict__ si, int ds,
                  const float* __restrict__ tu, const float* __restrict__ ti, int dt, const float* __restrict__ A,
                  const float* __restrict__ a, float scale /* weight / (B * width_out) */, float* __restrict__ gu,
                  float* __restrict__ gi, double* __restrict__ loss_accum) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int win = kind ? 2 * ds : ds, wout = kind ? 2 * dt : dt;
  float* xs = sm + warp * (win + wout);   // student feature
  float* r = xs + win;                    // residual (adapted student - teacher)
  float ls = 0.f;
  const int64_t nw = (int64_t)gridDim.x * kFeatWarps;
  for (int64_t b = (int64_t)blockIdx.x * kFeatWarps + warp; b < B; b += nw) {
    const int64_t u = user[b], it = item[b];
    if (u < 0 || u >= U || it < 0 || it >= I) continue;   // the step kernel reports the bad index
    const float *psu = su + u * ds, *psi = si + it * ds, *ptu = tu + u * dt, *pti = ti + it * dt;
    for (int k = lane; k < win; k += 32) xs[k] = kind ? (k < ds ? psu[k] : psi[k - ds]) : psu[k] * psi[k];
    __syncwarp();
    for (int j = lane; j < wout; j += 32) {
      const float t = kind ? (j < dt ? ptu[j] : pti[j - dt]) : ptu[j] * pti[j];
      float v;
      if (A != nullptr) {
        v = a[j];
        const float* row = A + (int64_t)j * win;
        for (int k = 0; k < win; ++k) v = fmaf(row[k], xs[k], v);
      } else {
        v = xs[j];
      }
      const float d = v - t;
      r[j] = d;
      ls += d * d;
    }
    __syncwarp();
    // gradient wrt the student feature: g_k = 2 * scale * sum_j A[j][k] r_j, scattered through the feature's definition
    for (int k = lane; k < win; k += 32) {
      float g;
      if (A != nullptr) {
        g = 0.f;
        for (int j = 0; j < wout; ++j) g = fmaf(A[(int64_t)j * win + k], r[j], g);
      } else {
        g = r[k];
      }
      g *= 2.f * scale;
      if (kind) {
        if (k < ds) atomicAdd(gu + u * ds + k, g);
        else atomicAdd(gi + it * ds + (k - ds), g);
      } else {
        atomicAdd(gu + u * ds + k, g * psi[k]);
        atomicAdd(gi + it * ds + k, g * psu[k]);
      }
    }
    __syncwarp();
  }
  ls = warp_sum(ls);
  if (lane == 0 && loss_accum != nullptr && ls != 0.f) atomicAdd(loss_accum, (double)ls * (double)scale);
}

}  // namespace

extern "C" int ncf_loss_grad_kd(const float* logits, const float* label, const float* teacher_logits, float w_task,
                                float w_kd, float temperature, int32_t kd_mode, int64_t B, double* loss_accum,
                                float* dlogit, void* stream) {
  NCF_REQUIRE(B > 0, "ncf_loss_grad_kd: empty batch");
  NCF_REQUIRE(logits && label, "ncf_loss_grad_kd: null pointer");
  NCF_REQUIRE(kd_mode == 0 || kd_mode == 1, "ncf_loss_grad_kd: kd_mode must be 0 (logit MSE) or 1 (soft targets)");
  NCF_REQUIRE(kd_mode == 0 || temperature > 0.f, "ncf_loss_grad_kd: temperature must be positive");
  int64_t blocks = (B + 255) / 256;
  if (blocks > 4 * ncf::num_sms()) blocks = 4 * ncf::num_sms();
  loss_grad_kd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, label, teacher_logits, w_task, w_kd, temperature,
                                                                    kd_mode, B, 1.f / (float)B, loss_accum, dlogit);
  NCF_LAUNCH_CHECK("loss_grad_kd_kernel");
  return NCF_OK;
}

extern "C" int ncf_feature_kd(const NcfModel* student, const NcfModel* teacher, const NcfGrads* g, const int64_t* user,
                              const int64_t* item, int64_t B, int32_t kind, const float* adapter_w, const float* adapter_b,
                              float weight, double* loss_accum, void* stream) {
  int rc = ncf::validate_model(student);
  if (rc != NCF_OK) return rc;
  if ((rc = ncf::validate_model(teacher)) != NCF_OK) return rc;
  NCF_REQUIRE(B > 0 && g && user && item, "ncf_feature_kd: empty batch or null pointer");
  NCF_REQUIRE(kind == 0 || kind == 1, "ncf_feature_kd: kind must be 0 (gmf_features) or 1 (mlp_input)");
  NCF_REQUIRE(student->user_num == teacher->user_num && student->item_num == teacher->item_num,
              "ncf_feature_kd: teacher and student index different tables");
  const bool need_gmf = kind == 0;
  NCF_REQUIRE(need_gmf ? (student->model_type != NCF_MLP && teacher->model_type != NCF_MLP)
                       : (student->model_type != NCF_GMF && teacher->model_type != NCF_GMF),
              "ncf_feature_kd: both models need the %s tables", need_gmf ? "GMF" : "MLP");
  const int ds = need_gmf ? student->factor_num : student->mlp_dim;
  const int dt = need_gmf ? teacher->factor_num : teacher->mlp_dim;
  NCF_REQUIRE((adapter_w != nullptr) == (adapter_b != nullptr), "ncf_feature_kd: adapter weight and bias go together");
  NCF_REQUIRE(adapter_w != nullptr || ds == dt, "ncf_feature_kd: widths differ (%d vs %d) and no adapter was given", ds, dt);
  const int win = kind ? 2 * ds : ds, wout = kind ? 2 * dt : dt;
  NCF_REQUIRE(win + wout <= kMaxFeat * 2 && (size_t)(win + wout) * 4 * kFeatWarps <= 200 * 1024,
              "ncf_feature_kd: feature widths %d + %d exceed the shared-memory budget", win, wout);
  float* gu = need_gmf ? g->g_user_gmf : g->g_user_mlp;
  float* gi = need_gmf ? g->g_item_gmf : g->g_item_mlp;
  NCF_REQUIRE(gu && gi, "ncf_feature_kd: gradient buffers are NULL");
  const size_t smem = (size_t)(win + wout) * 4 * kFeatWarps;
  if (smem > 48 * 1024)
    NCF_CUDA(cudaFuncSetAttribute(feature_kd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (B + kFeatWarps - 1) / kFeatWarps;
  if (blocks > 8 * ncf::num_sms()) blocks = 8 * ncf::num_sms();
  feature_kd_kernel<<<(int)blocks, kFeatWarps * 32, smem, (cudaStream_t)stream>>>(
      user, item, B, student->user_num, student->item_num, kind, need_gmf ? student->embed_user_gmf : student->embed_user_mlp,
      need_gmf ? student->embed_item_gmf : student->embed_item_mlp, ds, need_gmf ? teacher->embed_user_gmf : teacher->embed_user_mlp,
      need_gmf ? teacher->embed_item_gmf : teacher->embed_item_mlp, dt, adapter_w, adapter_b,
      weight / ((float)B * (float)wout), gu, gi, loss_accum);
  NCF_LAUNCH_CHECK("feature_kd_kernel");
  return NCF_OK;
}
