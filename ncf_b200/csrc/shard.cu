// Row-sharded tables (BASELINE config 5): the device-side pieces of the embedding-row exchange.
// Row r of a table lives on rank r % world at local index r / world.  Per step a rank (i) buckets
// its samples by the owner of their item row, (ii) asks the owners for the rows, (iii) runs the
// fused step on its local user rows + the received item rows, (iv) returns the item-row gradients
// to the owners.  The collectives themselves are NCCL all-to-alls issued by the host
// (ncf_b200/dist.py); the kernels here pack, unpack and scatter-add.  HBM-bound byte shuffling.
#include "common.cuh"

namespace {

constexpr int kMaxWorld = 64;

__global__ void owner_count_kernel(const int64_t* __restrict__ item, int64_t n, int world,
                                   int32_t* __restrict__ counts) {
  __shared__ int32_t h[kMaxWorld];
  for (int i = threadIdx.x; i < world; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n;
       b += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&h[(int)(item[b] % world)], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < world; i += blockDim.x)
    if (h[i]) atomicAdd(&counts[i], h[i]);
}

// counts[0..world) -> cursor[0..world) = exclusive prefix (single thread; world <= 64)
__global__ void owner_scan_kernel(const int32_t* __restrict__ counts, int world,
                                  int32_t* __restrict__ cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int32_t run = 0;
    for (int i = 0; i < world; ++i) { cursor[i] = run; run += counts[i]; }
  }
}

// perm[pos] = sample index, grouped by owner (order inside a group is arbitrary);
// local_idx[pos] = item / world of that sample.
__global__ void owner_place_kernel(const int64_t* __restrict__ item, int64_t n, int world,
                                   int32_t* __restrict__ cursor, int64_t* __restrict__ perm,
                                   int64_t* __restrict__ local_idx) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n;
       b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t it = item[b];
    const int pos = atomicAdd(&cursor[(int)(it % world)], 1);
    perm[pos] = b;
    local_idx[pos] = it / world;
  }
}

// out[i][:] = table[idx[i]][:]   (one warp per row, 16-byte pieces when dim % 4 == 0)
__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                   int64_t n, int dim, int64_t rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < n; i += nw) {
    const int64_t r = idx[i];
    const bool ok = r >= 0 && r < rows;
    const float* src = table + (ok ? r : 0) * dim;
    float* dst = out + i * dim;
    if ((dim & 3) == 0) {
      for (int c = lane * 4; c < dim; c += 128)
        *reinterpret_cast<float4*>(dst + c) = ok ? ldg4(src + c) : make_float4(0, 0, 0, 0);
    } else {
      for (int c = lane; c < dim; c += 32) dst[c] = ok ? __ldg(src + c) : 0.f;
    }
  }
}

// table[idx[i]][:] += in[i][:]   (vector REDs; duplicates in idx accumulate)
__global__ void scatter_add_rows_kernel(float* __restrict__ table, const int64_t* __restrict__ idx,
                                        int64_t n, int dim, int64_t rows,
                                        const float* __restrict__ in) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < n; i += nw) {
    const int64_t r = idx[i];
    if (r < 0 || r >= rows) continue;
    float* dst = table + r * dim;
    const float* src = in + i * dim;
    if ((dim & 3) == 0) {
      for (int c = lane * 4; c < dim; c += 128) red_add4(dst + c, ldg4(src + c));
    } else {
      for (int c = lane; c < dim; c += 32) atomicAdd(dst + c, __ldg(src + c));
    }
  }
}

// out[i] = in[perm[i]]  for int64 / float payloads
template <typename T>
__global__ void permute_kernel(const T* __restrict__ in, const int64_t* __restrict__ perm, int64_t n,
                               T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[perm[i]];
}

inline int blocks_for(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)ncf::num_sms() * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int ncf_bucket_by_owner(const int64_t* item, int64_t n, int32_t world, int64_t* perm,
                                   int64_t* local_idx, int32_t* counts, int32_t* cursor,
                                   void* stream) {
  NCF_REQUIRE(n >= 0 && world >= 1 && world <= kMaxWorld, "ncf_bucket_by_owner: bad n/world");
  NCF_REQUIRE(counts && cursor, "ncf_bucket_by_owner: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  NCF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * world, st));
  if (n > 0) {
    NCF_REQUIRE(item && perm && local_idx, "ncf_bucket_by_owner: null pointer");
    owner_count_kernel<<<blocks_for(n, 256), 256, 0, st>>>(item, n, world, counts);
    NCF_LAUNCH_CHECK("owner_count_kernel");
  }
  owner_scan_kernel<<<1, 32, 0, st>>>(counts, world, cursor);
  NCF_LAUNCH_CHECK("owner_scan_kernel");
  if (n > 0) {
    owner_place_kernel<<<blocks_for(n, 256), 256, 0, st>>>(item, n, world, cursor, perm, local_idx);
    NCF_LAUNCH_CHECK("owner_place_kernel");
  }
  return NCF_OK;
}

extern "C" int ncf_gather_rows(const float* table, const int64_t* idx, int64_t n, int32_t dim,
                               int64_t rows, float* out, void* stream) {
  NCF_REQUIRE(n >= 0 && dim > 0 && rows > 0, "ncf_gather_rows: bad sizes");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(table && idx && out, "ncf_gather_rows: null pointer");
  gather_rows_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(table, idx, n, dim, rows, out);
  NCF_LAUNCH_CHECK("gather_rows_kernel");
  return NCF_OK;
}

extern "C" int ncf_scatter_add_rows(float* table, const int64_t* idx, int64_t n, int32_t dim,
                                    int64_t rows, const float* in, void* stream) {
  NCF_REQUIRE(n >= 0 && dim > 0 && rows > 0, "ncf_scatter_add_rows: bad sizes");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(table && idx && in, "ncf_scatter_add_rows: null pointer");
  scatter_add_rows_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(table, idx, n, dim, rows, in);
  NCF_LAUNCH_CHECK("scatter_add_rows_kernel");
  return NCF_OK;
}

extern "C" int ncf_permute_i64(const int64_t* in, const int64_t* perm, int64_t n, int64_t* out,
                               void* stream) {
  NCF_REQUIRE(n >= 0, "ncf_permute_i64: negative n");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(in && perm && out, "ncf_permute_i64: null pointer");
  permute_kernel<int64_t><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, perm, n, out);
  NCF_LAUNCH_CHECK("permute_kernel<i64>");
  return NCF_OK;
}

extern "C" int ncf_permute_f32(const float* in, const int64_t* perm, int64_t n, float* out,
                               void* stream) {
  NCF_REQUIRE(n >= 0, "ncf_permute_f32: negative n");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(in && perm && out, "ncf_permute_f32: null pointer");
  permute_kernel<float><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, perm, n, out);
  NCF_LAUNCH_CHECK("permute_kernel<f32>");
  return NCF_OK;
}
