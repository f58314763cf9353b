// a1 CSR of observed pairs, a2 Philox negative sampler with CSR rejection, a3 epoch shuffle.
// Replaces reference src/data/datasets.py:20-24 (dok fill), :53-69 (ng_sample) and the
// DataLoader(shuffle=True) batching of scripts/train_neumf.py:55.  HBM-bound integer work.
#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr int kMaxAttempts = 1 << 16;  // a user who interacted with every item cannot loop forever

__global__ void csr_count_kernel(const int64_t* __restrict__ pos_user, const int64_t* __restrict__ pos_item, int64_t P,
                                 int64_t U, unsigned long long* __restrict__ count, int* __restrict__ bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < P; i += stride) {
    int64_t u = pos_user[i];
    const int64_t it = pos_item[i];
    if (it < 0 || it > 0x7fffffffll) *bad = 1;   // columns are int32
    if (u < 0 || u >= U) { *bad = 1; continue; }  // dropped from the CSR and reported through bad_flag
    atomicAdd(&count[u], 1ull);
  }
}

// Single-block exclusive scan of count[0..n) into rowptr[0..n]; also copies it into cursor.
__global__ void csr_scan_kernel(const unsigned long long* __restrict__ count, int64_t n,
                                int64_t* __restrict__ rowptr,
                                unsigned long long* __restrict__ cursor) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const int64_t lo = min((int64_t)t * chunk, n), hi = min(lo + chunk, n);
  unsigned long long s = 0;
  for (int64_t i = lo; i < hi; ++i) s += count[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { unsigned long long v = part[i]; part[i] = run; run += v; }
    rowptr[n] = (int64_t)run;
  }
  __syncthreads();
  unsigned long long run = part[t];
  for (int64_t i = lo; i < hi; ++i) {
    rowptr[i] = (int64_t)run;
    cursor[i] = run;
    run += count[i];
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ pos_user,
                                const int64_t* __restrict__ pos_item, int64_t P, int64_t U,
                                unsigned long long* __restrict__ cursor, int32_t* __restrict__ tmp) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < P; i += stride) {
    int64_t u = pos_user[i];
    if (u < 0 || u >= U) continue;
    unsigned long long slot = atomicAdd(&cursor[u], 1ull);
    tmp[slot] = (int32_t)pos_item[i];
  }
}

// ---- per-row sort of the columns ------------------------------------------------------------------------
// csr_fill leaves every row's items in arrival order.  Rows are sorted in shared memory by a bitonic
// network over the next power of two (padding = INT32_MAX, never written back): one WARP per row up to
// kWarpRow items, one CTA per row up to kCtaRow items (128 KB).  Longer rows (more than 32 768 items of
// one user) fall back to the rank sort below.  Duplicate pairs are kept.
constexpr int kWarpRow = 256;
constexpr int kCtaRow = 32768;

template <bool WARP>
__device__ __forceinline__ void bitonic_sort_shared(int32_t* a, int N, int tid, int nthreads) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < N; i += nthreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const int32_t x = a[i], y = a[ixj];
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      if (WARP) __syncwarp(); else __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) csr_sort_rows_warp_kernel(const int64_t* __restrict__ rowptr, int64_t U,
                                                                 const int32_t* __restrict__ tmp, int32_t* __restrict__ col) {
  __shared__ int32_t buf[8][kWarpRow];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* a = buf[warp];
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t u = (int64_t)blockIdx.x * 8 + warp; u < U; u += nw) {
    const int64_t b = rowptr[u];
    const int64_t n = rowptr[u + 1] - b;
    if (n <= 0 || n > kWarpRow) continue;
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = lane; i < N; i += 32) a[i] = i < n ? tmp[b + i] : 0x7fffffff;
    __syncwarp();
    bitonic_sort_shared<true>(a, N, lane, 32);
    for (int i = lane; i < n; i += 32) col[b + i] = a[i];
    __syncwarp();
  }
}

// rows with lo < n <= hi items; launched once per size class so that the many medium rows run on CTAs with a
// small shared-memory footprint (several per SM) and only the few very long ones take a 128 KB CTA
__global__ void __launch_bounds__(256) csr_sort_rows_cta_kernel(const int64_t* __restrict__ rowptr, int64_t U,
                                                                const int32_t* __restrict__ tmp, int32_t* __restrict__ col,
                                                                int lo, int hi) {
  extern __shared__ int32_t big[];
  for (int64_t u = blockIdx.x; u < U; u += gridDim.x) {
    const int64_t b = rowptr[u];
    const int64_t n = rowptr[u + 1] - b;
    if (n <= lo || n > hi) continue;   // uniform over the CTA
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) big[i] = i < n ? tmp[b + i] : 0x7fffffff;
    __syncthreads();
    bitonic_sort_shared<false>(big, N, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < n; i += blockDim.x) col[b + i] = big[i];
    __syncthreads();
  }
}

// Rank sort inside a row (only rows longer than kCtaRow): position of an item = number of smaller items
// in its row.  Equal items (duplicate pairs) all write the same run of slots, so no slot is left unwritten.
__global__ void csr_rank_kernel(const int64_t* __restrict__ pos_user,
                                const int64_t* __restrict__ pos_item, int64_t P, int64_t U,
                                const int64_t* __restrict__ rowptr, const int32_t* __restrict__ tmp,
                                int32_t* __restrict__ col) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < P; i += stride) {
    int64_t u = pos_user[i];
    if (u < 0 || u >= U) continue;
    const int32_t it = (int32_t)pos_item[i];
    const int64_t b = rowptr[u], e = rowptr[u + 1];
    if (e - b <= kCtaRow) continue;
    int64_t less = 0, equal = 0;
    for (int64_t j = b; j < e; ++j) {
      int32_t x = __ldg(&tmp[j]);
      less += (x < it);
      equal += (x == it);
    }
    for (int64_t j = 0; j < equal; ++j) col[b + less + j] = it;
  }
}

__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ col, int64_t lo, int64_t hi,
                                             int32_t key) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t v = __ldg(&col[mid]);
    if (v == key) return true;
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Out-of-range users and users whose every draw was rejected kMaxAttempts times (they interacted with
// (nearly) every item; the reference loops forever there) get -1, which the training kernels report as
// a bad index (NaN logit) instead of silently training on an observed pair.
__global__ void sample_neg_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  const int64_t* __restrict__ pos_user, int64_t P, int64_t p_offset, int64_t user_num,
                                  int num_ng, uint32_t item_num, uint32_t seed_lo, uint32_t seed_hi,
                                  uint32_t epoch, int64_t* __restrict__ out) {
  const int64_t total = P * num_ng;
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; g < total; g += stride) {
    const int64_t p = g / num_ng;
    const int64_t u = pos_user[p];
    if (u < 0 || u >= user_num) { out[g] = -1; continue; }
    const int64_t lo = rowptr[u], hi = rowptr[u + 1];
    const uint64_t sid = (uint64_t)(p_offset * num_ng + g);
    int32_t j = -1;
    Philox4 r = {0, 0, 0, 0};
    int a = 0;
    for (; a < kMaxAttempts; ++a) {
      if ((a & 3) == 0)
        r = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)(a >> 2), epoch, seed_lo,
                          seed_hi);
      const uint32_t w = (a & 3) == 0 ? r.x : (a & 3) == 1 ? r.y : (a & 3) == 2 ? r.z : r.w;
      j = (int32_t)__umulhi(w, item_num);
      if (!csr_contains(col, lo, hi, j)) break;
    }
    out[g] = (a < kMaxAttempts) ? (int64_t)j : -1;
  }
}

__global__ void shuffle_epoch_kernel(const int64_t* __restrict__ pos_user,
                                     const int64_t* __restrict__ pos_item,
                                     const int64_t* __restrict__ neg_item, int64_t P, int num_ng,
                                     ShufflePerm perm, int64_t q_begin, int64_t count,
                                     int64_t* __restrict__ out_user, int64_t* __restrict__ out_item,
                                     float* __restrict__ out_label) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    const uint64_t s = shuffle_perm_apply(perm, (uint64_t)(q_begin + i));
    int64_t u, it;
    float y;
    if ((int64_t)s < P) {
      u = pos_user[s];
      it = pos_item[s];
      y = 1.0f;
    } else {
      const int64_t n = (int64_t)s - P;
      u = pos_user[n / num_ng];
      it = neg_item[n];
      y = 0.0f;
    }
    out_user[i] = u;
    out_item[i] = it;
    out_label[i] = y;
  }
}

inline int grid_for(int64_t n, int threads) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ncf::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int64_t ncf_csr_workspace_bytes(int64_t P, int64_t user_num) {
  // count[U] + cursor[U] (u64) + tmp[P] (i32) + bad flag
  return ncf::align_up(user_num * 8, 256) * 2 + ncf::align_up(P * 4, 256) + 256;
}

extern "C" int ncf_csr_build(const int64_t* pos_user, const int64_t* pos_item, int64_t P,
                             int64_t user_num, int64_t* rowptr, int32_t* col, int32_t* bad_flag, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(P >= 0 && user_num > 0, "ncf_csr_build: bad sizes P=%lld U=%lld", (long long)P,
              (long long)user_num);
  NCF_REQUIRE(rowptr && (P == 0 || (pos_user && pos_item && col)), "ncf_csr_build: null pointer");
  if (workspace_bytes < ncf_csr_workspace_bytes(P, user_num) || !workspace) {
    ncf::set_error("ncf_csr_build: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                   (long long)ncf_csr_workspace_bytes(P, user_num));
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  auto* count = (unsigned long long*)ws;
  ws += ncf::align_up(user_num * 8, 256);
  auto* cursor = (unsigned long long*)ws;
  ws += ncf::align_up(user_num * 8, 256);
  auto* tmp = (int32_t*)ws;
  ws += ncf::align_up(P * 4, 256);
  int* bad = (int*)ws;
  NCF_CUDA(cudaMemsetAsync(count, 0, user_num * 8, st));
  NCF_CUDA(cudaMemsetAsync(bad, 0, 4, st));
  const int T = 256;
  if (P > 0) {
    csr_count_kernel<<<grid_for(P, T), T, 0, st>>>(pos_user, pos_item, P, user_num, count, bad);
    NCF_LAUNCH_CHECK("csr_count");
  }
  csr_scan_kernel<<<1, 1024, 0, st>>>(count, user_num, rowptr, cursor);
  NCF_LAUNCH_CHECK("csr_scan");
  if (P > 0) {
    csr_fill_kernel<<<grid_for(P, T), T, 0, st>>>(pos_user, pos_item, P, user_num, cursor, tmp);
    NCF_LAUNCH_CHECK("csr_fill");
    csr_sort_rows_warp_kernel<<<grid_for(user_num * 32, T), T, 0, st>>>(rowptr, user_num, tmp, col);
    NCF_LAUNCH_CHECK("csr_sort_rows_warp");
    static bool attr_set = false;
    if (!attr_set) {
      NCF_CUDA(cudaFuncSetAttribute(csr_sort_rows_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtaRow * 4));
      attr_set = true;
    }
    constexpr int kMidRow = 4096;   // 16 KB of shared memory: 8+ CTAs per SM
    csr_sort_rows_cta_kernel<<<ncf::num_sms() * 8, T, kMidRow * 4, st>>>(rowptr, user_num, tmp, col, kWarpRow, kMidRow);
    NCF_LAUNCH_CHECK("csr_sort_rows_cta(mid)");
    csr_sort_rows_cta_kernel<<<ncf::num_sms(), T, kCtaRow * 4, st>>>(rowptr, user_num, tmp, col, kMidRow, kCtaRow);
    NCF_LAUNCH_CHECK("csr_sort_rows_cta(long)");
    csr_rank_kernel<<<grid_for(P, T), T, 0, st>>>(pos_user, pos_item, P, user_num, rowptr, tmp, col);
    NCF_LAUNCH_CHECK("csr_rank");
  }
  if (bad_flag != nullptr) NCF_CUDA(cudaMemcpyAsync(bad_flag, bad, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return NCF_OK;
}

extern "C" int ncf_sample_neg(const int64_t* rowptr, const int32_t* col, const int64_t* pos_user,
                              int64_t P, int64_t p_offset, int64_t user_num, int32_t num_ng, int64_t item_num,
                              uint64_t seed, uint64_t epoch, int64_t* out_neg_item, void* stream) {
  NCF_REQUIRE(P >= 0 && num_ng >= 0 && p_offset >= 0 && user_num > 0, "ncf_sample_neg: bad sizes");
  NCF_REQUIRE(item_num > 0 && item_num <= 0x7fffffffLL, "ncf_sample_neg: item_num %lld out of range",
              (long long)item_num);
  if (P == 0 || num_ng == 0) return NCF_OK;
  NCF_REQUIRE(rowptr && col && pos_user && out_neg_item, "ncf_sample_neg: null pointer");
  const int T = 256;
  sample_neg_kernel<<<grid_for(P * num_ng, T), T, 0, (cudaStream_t)stream>>>(
      rowptr, col, pos_user, P, p_offset, user_num, num_ng, (uint32_t)item_num, (uint32_t)seed,
      (uint32_t)(seed >> 32), (uint32_t)epoch, out_neg_item);
  NCF_LAUNCH_CHECK("sample_neg");
  return NCF_OK;
}

extern "C" int ncf_shuffle_epoch(const int64_t* pos_user, const int64_t* pos_item,
                                 const int64_t* neg_item, int64_t P, int32_t num_ng, uint64_t seed,
                                 uint64_t epoch, int64_t q_begin, int64_t count, int64_t* out_user,
                                 int64_t* out_item, float* out_label, void* stream) {
  NCF_REQUIRE(P >= 0 && num_ng >= 0 && q_begin >= 0 && count >= 0, "ncf_shuffle_epoch: bad sizes");
  const int64_t S = P * (1 + (int64_t)num_ng);
  NCF_REQUIRE(q_begin + count <= S, "ncf_shuffle_epoch: range [%lld,%lld) exceeds S=%lld",
              (long long)q_begin, (long long)(q_begin + count), (long long)S);
  if (count == 0) return NCF_OK;
  NCF_REQUIRE(pos_user && pos_item && (num_ng == 0 || neg_item) && out_user && out_item && out_label,
              "ncf_shuffle_epoch: null pointer");
  ShufflePerm perm = make_shuffle_perm((uint64_t)S, seed, epoch);
  const int T = 256;
  shuffle_epoch_kernel<<<grid_for(count, T), T, 0, (cudaStream_t)stream>>>(
      pos_user, pos_item, neg_item, P, num_ng, perm, q_begin, count, out_user, out_item, out_label);
  NCF_LAUNCH_CHECK("shuffle_epoch");
  return NCF_OK;
}
