// Library-level entry points and shared host helpers of the C ABI (include/ncf_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace ncf {

static thread_local char g_err[512] = "";
thread_local int g_tile_path = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NCF_OK;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return NCF_ERR_CUDA;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

int validate_model(const NcfModel* m) {
  NCF_REQUIRE(m != nullptr, "model is NULL");
  NCF_REQUIRE(m->model_type >= NCF_GMF && m->model_type <= NCF_NEUMF, "bad model_type %d",
              m->model_type);
  NCF_REQUIRE(m->factor_num > 0, "bad factor_num %d", m->factor_num);
  NCF_REQUIRE(m->num_layers >= 1 && m->num_layers <= NCF_MAX_LAYERS, "bad num_layers %d",
              m->num_layers);
  NCF_REQUIRE(m->mlp_dim == (m->factor_num << (m->num_layers - 1)),
              "mlp_dim %d != factor_num*2^(L-1)", m->mlp_dim);
  NCF_REQUIRE(m->user_num > 0 && m->item_num > 0, "bad table sizes");
  const bool gmf = m->model_type != NCF_MLP, mlp = m->model_type != NCF_GMF;
  if (gmf) NCF_REQUIRE(m->embed_user_gmf && m->embed_item_gmf, "GMF tables are NULL");
  if (mlp) {
    NCF_REQUIRE(m->embed_user_mlp && m->embed_item_mlp, "MLP tables are NULL");
    for (int k = 0; k < m->num_layers; ++k)
      NCF_REQUIRE(m->mlp_w[k] && m->mlp_b[k], "tower layer %d is NULL", k);
  }
  NCF_REQUIRE(m->predict_w && m->predict_b, "predict layer is NULL");
  return NCF_OK;
}

}  // namespace ncf

extern "C" int ncf_version(void) { return NCF_ABI_VERSION; }

extern "C" const char* ncf_last_error(void) { return ncf::g_err; }

extern "C" int ncf_last_tile_path(void) { return ncf::g_tile_path; }

namespace ncf {
static thread_local cudaEvent_t g_embed_event = nullptr;
int mark_embedding_grads_done(cudaStream_t st) {
  if (g_embed_event == nullptr) NCF_CUDA(cudaEventCreateWithFlags(&g_embed_event, cudaEventDisableTiming));
  NCF_CUDA(cudaEventRecord(g_embed_event, st));
  return NCF_OK;
}
}  // namespace ncf

extern "C" int ncf_wait_embedding_grads(void* stream) {
  NCF_REQUIRE(ncf::g_embed_event != nullptr, "ncf_wait_embedding_grads: no training step has run on this thread");
  NCF_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ncf::g_embed_event, 0));
  return NCF_OK;
}

extern "C" int64_t ncf_tower_param_count(int32_t model_type, int32_t factor_num,
                                         int32_t num_layers) {
  if (factor_num <= 0 || num_layers < 1 || num_layers > NCF_MAX_LAYERS) return -1;
  return ncf::make_tower_shape(model_type, factor_num, num_layers).total;
}
