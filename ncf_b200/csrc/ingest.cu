// Data-side kernels around the hot path: the reference's on-disk formats and its
// leave-one-out preprocessing, as byte / integer work on the device.
//
//   ncf_text_line_starts / ncf_text_parse_ints
//       u.train.rating (`user<TAB>item` per line) and u.test.negative (`(user,pos)<TAB>neg1<TAB>...`), written by
//       reference src/data/preprocessing.py:137-154 and read by load_all (src/data/datasets.py:9-36: a pandas
//       read plus a Python loop that eval()s every line).  Here: byte-parallel line index (two passes around
//       an exclusive scan), then one thread per line takes the first K integers (maximal digit runs).
//   ncf_leave_one_out_split
//       LeaveOneOutPreprocessor.temporal_split (preprocessing.py:45-90): bucket the ratings by user, sort
//       every user's row by (timestamp, position in the file) in shared memory, last = test for users with
//       at least two interactions, the rest = train in that order.
//   ncf_eval_negatives
//       generate_test_negatives (preprocessing.py:92-135): up to K distinct items per test user, uniform
//       over [0, num_items), none of the user's own items, at most 10 K draws, ascending.  One warp per
//       user; draws are Philox4x32-10 words (the reference's numpy MT19937 stream cannot be reproduced).
// HBM / latency-bound integer work; one-off per data set.
#include <algorithm>

#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr int kT = 256;
constexpr int kBytesPerThread = 16;
constexpr int kChunk = kT * kBytesPerThread;   // bytes per CTA

__device__ __forceinline__ bool is_eol(uint8_t c) { return c == '\n' || c == '\r'; }

// line start: first byte of the text or a byte after '\n', that is not itself an end-of-line byte
__device__ __forceinline__ bool starts_line(const uint8_t* __restrict__ text, int64_t i) {
  return !is_eol(text[i]) && (i == 0 || text[i - 1] == '\n');
}

// pass 1 (out == nullptr): block_count[b] = line starts in the CTA's chunk.
// pass 2: out[block_off[b] + rank within the chunk] = position.
__global__ void __launch_bounds__(kT) line_starts_kernel(const uint8_t* __restrict__ text, int64_t n,
                                                         int64_t* __restrict__ block_count,
                                                         const int64_t* __restrict__ block_off, int64_t* __restrict__ out,
                                                         int64_t cap) {
  __shared__ int warp_sum_sm[kT / 32];
  const int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kBytesPerThread;
  uint32_t mask = 0;   // bit j: byte base + j starts a line
  for (int j = 0; j < kBytesPerThread; ++j)
    if (base + j < n && starts_line(text, base + j)) mask |= 1u << j;
  const int mine = __popc(mask);
  // exclusive prefix of `mine` over the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum_sm[warp] = incl;
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < kT / 32; ++w) {
    if (w < warp) before += warp_sum_sm[w];
    total += warp_sum_sm[w];
  }
  if (out == nullptr) {
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
    return;
  }
  int64_t pos = block_off[blockIdx.x] + before + incl - mine;
  for (int j = 0; j < kBytesPerThread; ++j)
    if (mask >> j & 1u) {
      if (pos < cap) out[pos] = base + j;
      ++pos;
    }
}

// single CTA: exclusive scan of count[0..n) -> off[0..n), total -> *total_out
__global__ void scan_i64_kernel(const int64_t* __restrict__ count, int64_t n, int64_t* __restrict__ off,
                                int64_t* __restrict__ total_out) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const int64_t lo = min((int64_t)t * chunk, n), hi = min(lo + chunk, n);
  int64_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += count[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    int64_t run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { const int64_t v = part[i]; part[i] = run; run += v; }
    if (total_out) *total_out = run;
  }
  __syncthreads();
  int64_t run = part[t];
  for (int64_t i = lo; i < hi; ++i) { const int64_t v = count[i]; off[i] = run; run += v; }
}

// one thread per line: the first K integers (maximal digit runs, optional leading '-') of the line.
// status bit 0: some line had fewer than K integers (the missing ones are -1); bit 1 (exact only): more than K.
__global__ void parse_ints_kernel(const uint8_t* __restrict__ text, int64_t n, const int64_t* __restrict__ line_start,
                                  int64_t n_lines, int K, int exact, int64_t* __restrict__ out, int* __restrict__ status) {
  for (int64_t ln = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ln < n_lines; ln += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = line_start[ln];
    int found = 0;
    bool in_num = false, neg = false;
    int64_t v = 0;
    for (; i <= n; ++i) {
      const uint8_t c = (i < n) ? text[i] : (uint8_t)'\n';
      const bool dig = c >= '0' && c <= '9';
      if (dig) {
        if (!in_num) { in_num = true; v = 0; neg = (i > line_start[ln] && text[i - 1] == '-'); }
        v = v * 10 + (c - '0');
      } else if (in_num) {
        if (found < K) out[ln * K + found] = neg ? -v : v;
        ++found;
        in_num = false;
      }
      if (c == '\n') break;
    }
    for (int j = found; j < K; ++j) out[ln * K + j] = -1;
    if (found < K) atomicOr(status, 1);
    if (exact && found > K) atomicOr(status, 2);
  }
}

// ---- leave-one-out split ---------------------------------------------------------------------------------
__global__ void split_count_kernel(const int64_t* __restrict__ user, int64_t n, int64_t U,
                                   unsigned long long* __restrict__ count, int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = user[i];
    if (u < 0 || u >= U) { *bad = 1; continue; }
    atomicAdd(&count[u], 1ull);
  }
}

// rowptr (exclusive scan of count) and, per user, the number of users with a test row before it
__global__ void split_scan_kernel(const unsigned long long* __restrict__ count, int64_t U, int64_t* __restrict__ rowptr,
                                  unsigned long long* __restrict__ cursor, int64_t* __restrict__ tests_before,
                                  int64_t* __restrict__ totals /* [2]: n_train, n_test */) {
  __shared__ unsigned long long part[1024];
  __shared__ unsigned long long tpart[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (U + blockDim.x - 1) / blockDim.x;
  const int64_t lo = min((int64_t)t * chunk, U), hi = min(lo + chunk, U);
  unsigned long long s = 0, ts = 0;
  for (int64_t i = lo; i < hi; ++i) { s += count[i]; ts += count[i] >= 2; }
  part[t] = s;
  tpart[t] = ts;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0, trun = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) {
      const unsigned long long v = part[i], tv = tpart[i];
      part[i] = run; tpart[i] = trun;
      run += v; trun += tv;
    }
    rowptr[U] = (int64_t)run;
    totals[0] = (int64_t)(run - trun);
    totals[1] = (int64_t)trun;
  }
  __syncthreads();
  unsigned long long run = part[t], trun = tpart[t];
  for (int64_t i = lo; i < hi; ++i) {
    rowptr[i] = (int64_t)run;
    cursor[i] = run;
    tests_before[i] = (int64_t)trun;
    run += count[i];
    trun += count[i] >= 2;
  }
}

// key = (timestamp << 32) | position in the file: sorting the keys of a row orders it by time, ties by file order
__global__ void split_fill_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ ts, int64_t n, int64_t U,
                                  unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ keys,
                                  int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = user[i];
    if (u < 0 || u >= U) continue;
    const int64_t t = ts[i];
    if (t < 0 || t > 0xffffffffll) *bad = 1;   // timestamps are taken as 32-bit unsigned seconds
    keys[atomicAdd(&cursor[u], 1ull)] = ((unsigned long long)(t & 0xffffffffll) << 32) | (unsigned long long)i;
  }
}

constexpr int kSplitRow = 16384;  // keys of one user sorted by one CTA in shared memory (128 KB; ML-20M's busiest user: 9 254)

// one CTA per user row (rows of <= 32 keys: one warp does it, the others idle; rows beyond kSplitRow: flagged)
__global__ void __launch_bounds__(kT) split_sort_emit_kernel(const int64_t* __restrict__ rowptr,
                                                             const int64_t* __restrict__ tests_before, int64_t U,
                                                             unsigned long long* __restrict__ keys,
                                                             const int64_t* __restrict__ item, int64_t* __restrict__ train_user,
                                                             int64_t* __restrict__ train_item, int64_t* __restrict__ test_user,
                                                             int64_t* __restrict__ test_item, int* __restrict__ bad) {
  extern __shared__ unsigned long long row[];
  for (int64_t u = blockIdx.x; u < U; u += gridDim.x) {
    const int64_t b = rowptr[u];
    const int n = (int)min(rowptr[u + 1] - b, (int64_t)(kSplitRow + 1));
    if (n == 0) continue;
    if (n > kSplitRow) {
      if (threadIdx.x == 0) *bad = 2;
      continue;
    }
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = threadIdx.x; i < N; i += kT) row[i] = i < n ? keys[b + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < N; i += kT) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool up = (i & k) == 0;
            const unsigned long long x = row[i], y = row[ixj];
            if ((x > y) == up) { row[i] = y; row[ixj] = x; }
          }
        }
        __syncthreads();
      }
    const bool has_test = n >= 2;
    const int64_t tb = tests_before[u];
    const int64_t train0 = b - tb;            // one row left out for every earlier user that has a test row
    const int n_train = has_test ? n - 1 : n;
    for (int i = threadIdx.x; i < n_train; i += kT) {
      const int64_t src = (int64_t)(row[i] & 0xffffffffull);
      train_user[train0 + i] = u;
      train_item[train0 + i] = item[src];
    }
    if (has_test && threadIdx.x == 0) {
      test_user[tb] = u;
      test_item[tb] = item[(int64_t)(row[n - 1] & 0xffffffffull)];
    }
    __syncthreads();
  }
}

// ---- evaluation negatives ---------------------------------------------------------------------------------
constexpr uint32_t kNegsTag = 0x4E454753u;   // 'NEGS'
constexpr int kMaxNeg = 1024;

__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ col, int64_t lo, int64_t hi, int32_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(&col[mid]);
    if (v == key) return true;
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(kT) eval_negatives_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            const int64_t* __restrict__ test_user, int64_t n, int64_t U,
                                                            uint32_t num_items, int K, uint32_t seed_lo, uint32_t seed_hi,
                                                            int64_t* __restrict__ out, int32_t* __restrict__ count) {
  extern __shared__ int32_t got_sm[];            // [warps][Kpad]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int Kpad = 32;
  while (Kpad < K) Kpad <<= 1;
  int32_t* got = got_sm + warp * Kpad;
  const int64_t nw = (int64_t)gridDim.x * (kT / 32);
  for (int64_t r = (int64_t)blockIdx.x * (kT / 32) + warp; r < n; r += nw) {
    const int64_t u = test_user[r];
    const bool ok = u >= 0 && u < U;
    const int64_t lo = ok ? rowptr[u] : 0, hi = ok ? rowptr[u + 1] : 0;
    int have = 0;
    Philox4 w = {0, 0, 0, 0};
    for (int a = 0; a < 10 * K && have < K && ok; ++a) {     // uniform over the warp
      if ((a & 3) == 0) w = philox4x32_10((uint32_t)r, (uint32_t)((uint64_t)r >> 32), (uint32_t)(a >> 2), kNegsTag, seed_lo, seed_hi);
      const uint32_t word = (a & 3) == 0 ? w.x : (a & 3) == 1 ? w.y : (a & 3) == 2 ? w.z : w.w;
      const int32_t j = (int32_t)__umulhi(word, num_items);
      bool dup = false;
      for (int k = lane; k < have; k += 32) dup |= got[k] == j;
      dup = __any_sync(0xffffffffu, dup);
      if (dup) continue;
      bool mine = false;
      if (lane == 0) mine = row_contains(col, lo, hi, j);
      mine = __shfl_sync(0xffffffffu, (int)mine, 0) != 0;
      if (mine) continue;
      if (lane == 0) got[have] = j;
      ++have;
      __syncwarp();
    }
    for (int k = have + lane; k < Kpad; k += 32) got[k] = 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= Kpad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < Kpad; i += 32) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool up = (i & k) == 0;
            const int32_t x = got[i], y = got[ixj];
            if ((x > y) == up) { got[i] = y; got[ixj] = x; }
          }
        }
        __syncwarp();
      }
    for (int k = lane; k < K; k += 32) out[r * K + k] = k < have ? (int64_t)got[k] : -1;
    if (lane == 0) count[r] = have;
    __syncwarp();
  }
}

inline int grid_for(int64_t n, int per_block) {
  int64_t blocks = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)ncf::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

// workspace of ncf_text_line_starts: two int64 per 4 KB chunk of text
extern "C" int64_t ncf_text_workspace_bytes(int64_t nbytes) {
  if (nbytes < 0) return -1;
  const int64_t blocks = (nbytes + kChunk - 1) / kChunk;
  return ncf::align_up(blocks * 8, 256) * 2 + 256;
}

extern "C" int ncf_text_line_starts(const uint8_t* text, int64_t nbytes, int64_t* line_start, int64_t cap,
                                    int64_t* n_lines, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(nbytes >= 0 && n_lines != nullptr && cap >= 0, "ncf_text_line_starts: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (nbytes == 0) {
    NCF_CUDA(cudaMemsetAsync(n_lines, 0, 8, st));
    return NCF_OK;
  }
  NCF_REQUIRE(text != nullptr, "ncf_text_line_starts: text is NULL");
  if (!workspace || workspace_bytes < ncf_text_workspace_bytes(nbytes)) {
    ncf::set_error("ncf_text_line_starts: workspace too small");
    return NCF_ERR_WORKSPACE;
  }
  const int64_t blocks = (nbytes + kChunk - 1) / kChunk;
  NCF_REQUIRE(blocks < (1ll << 31), "ncf_text_line_starts: text too large");
  int64_t* block_count = (int64_t*)workspace;
  int64_t* block_off = (int64_t*)((char*)workspace + ncf::align_up(blocks * 8, 256));
  line_starts_kernel<<<(unsigned)blocks, kT, 0, st>>>(text, nbytes, block_count, nullptr, nullptr, 0);
  NCF_LAUNCH_CHECK("line_starts_kernel(count)");
  scan_i64_kernel<<<1, 1024, 0, st>>>(block_count, blocks, block_off, n_lines);
  NCF_LAUNCH_CHECK("scan_i64_kernel");
  if (line_start != nullptr && cap > 0) {
    line_starts_kernel<<<(unsigned)blocks, kT, 0, st>>>(text, nbytes, block_count, block_off, line_start, cap);
    NCF_LAUNCH_CHECK("line_starts_kernel(write)");
  }
  return NCF_OK;
}

extern "C" int ncf_text_parse_ints(const uint8_t* text, int64_t nbytes, const int64_t* line_start, int64_t n_lines,
                                   int32_t K, int32_t exact, int64_t* out, int32_t* status, void* stream) {
  NCF_REQUIRE(nbytes >= 0 && n_lines >= 0 && K >= 1, "ncf_text_parse_ints: bad argument");
  NCF_REQUIRE(status != nullptr, "ncf_text_parse_ints: status is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  NCF_CUDA(cudaMemsetAsync(status, 0, 4, st));
  if (n_lines == 0) return NCF_OK;
  NCF_REQUIRE(text && line_start && out, "ncf_text_parse_ints: null pointer");
  parse_ints_kernel<<<grid_for(n_lines, 128), 128, 0, st>>>(text, nbytes, line_start, n_lines, K, exact, out, status);
  NCF_LAUNCH_CHECK("parse_ints_kernel");
  return NCF_OK;
}

// count[U] + cursor[U] (u64) + rowptr[U+1] + tests_before[U] (i64) + keys[n] (u64) + bad flag
extern "C" int64_t ncf_split_workspace_bytes(int64_t n, int64_t user_num) {
  if (n < 0 || user_num <= 0) return -1;
  return ncf::align_up(user_num * 8, 256) * 3 + ncf::align_up((user_num + 1) * 8, 256) + ncf::align_up(n * 8, 256) + 256;
}

extern "C" int ncf_leave_one_out_split(const int64_t* user, const int64_t* item, const int64_t* timestamp, int64_t n,
                                       int64_t user_num, int64_t* train_user, int64_t* train_item, int64_t* test_user,
                                       int64_t* test_item, int64_t* totals, int32_t* bad_flag, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(n >= 0 && user_num > 0 && n < (1ll << 32), "ncf_leave_one_out_split: bad sizes");
  NCF_REQUIRE(totals != nullptr, "ncf_leave_one_out_split: totals is NULL");
  if (!workspace || workspace_bytes < ncf_split_workspace_bytes(n, user_num)) {
    ncf::set_error("ncf_leave_one_out_split: workspace too small");
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  auto* count = (unsigned long long*)ws;       ws += ncf::align_up(user_num * 8, 256);
  auto* cursor = (unsigned long long*)ws;      ws += ncf::align_up(user_num * 8, 256);
  auto* tests_before = (int64_t*)ws;           ws += ncf::align_up(user_num * 8, 256);
  auto* rowptr = (int64_t*)ws;                 ws += ncf::align_up((user_num + 1) * 8, 256);
  auto* keys = (unsigned long long*)ws;        ws += ncf::align_up(n * 8, 256);
  int* bad = (int*)ws;
  NCF_CUDA(cudaMemsetAsync(count, 0, user_num * 8, st));
  NCF_CUDA(cudaMemsetAsync(bad, 0, 4, st));
  if (n > 0) {
    NCF_REQUIRE(user && item && timestamp && train_user && train_item && test_user && test_item,
                "ncf_leave_one_out_split: null pointer");
    split_count_kernel<<<grid_for(n, kT), kT, 0, st>>>(user, n, user_num, count, bad);
    NCF_LAUNCH_CHECK("split_count_kernel");
  }
  split_scan_kernel<<<1, 1024, 0, st>>>(count, user_num, rowptr, cursor, tests_before, totals);
  NCF_LAUNCH_CHECK("split_scan_kernel");
  if (n > 0) {
    split_fill_kernel<<<grid_for(n, kT), kT, 0, st>>>(user, timestamp, n, user_num, cursor, keys, bad);
    NCF_LAUNCH_CHECK("split_fill_kernel");
    static bool attr_set = false;
    if (!attr_set) {
      NCF_CUDA(cudaFuncSetAttribute(split_sort_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSplitRow * 8));
      attr_set = true;
    }
    const int64_t blocks = std::min<int64_t>(user_num, (int64_t)ncf::num_sms() * 4);
    split_sort_emit_kernel<<<(unsigned)blocks, kT, kSplitRow * 8, st>>>(rowptr, tests_before, user_num, keys, item, train_user,
                                                                      train_item, test_user, test_item, bad);
    NCF_LAUNCH_CHECK("split_sort_emit_kernel");
  }
  if (bad_flag != nullptr) NCF_CUDA(cudaMemcpyAsync(bad_flag, bad, 4, cudaMemcpyDeviceToDevice, st));
  return NCF_OK;
}

extern "C" int ncf_eval_negatives(const int64_t* rowptr, const int32_t* col, const int64_t* test_user, int64_t n,
                                  int64_t user_num, int64_t num_items, int32_t K, uint64_t seed, int64_t* out,
                                  int32_t* count, void* stream) {
  NCF_REQUIRE(n >= 0 && user_num > 0 && num_items > 0 && num_items <= 0x7fffffffLL, "ncf_eval_negatives: bad sizes");
  NCF_REQUIRE(K >= 1 && K <= kMaxNeg, "ncf_eval_negatives: K=%d outside [1, %d]", K, kMaxNeg);
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(rowptr && col && test_user && out && count, "ncf_eval_negatives: null pointer");
  int Kpad = 32;
  while (Kpad < K) Kpad <<= 1;
  const size_t smem = (size_t)(kT / 32) * Kpad * 4;
  eval_negatives_kernel<<<grid_for(n, kT / 32), kT, smem, (cudaStream_t)stream>>>(
      rowptr, col, test_user, n, user_num, (uint32_t)num_items, K, (uint32_t)seed, (uint32_t)(seed >> 32), out, count);
  NCF_LAUNCH_CHECK("eval_negatives_kernel");
  return NCF_OK;
}
