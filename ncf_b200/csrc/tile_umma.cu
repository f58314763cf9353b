// tcgen05 / TMEM version of the tower arithmetic for large batches (sm_100a).
//
// The tower (reference: src/ncf/models.py:86-107, MLP_layers + predict_layer, and autograd's
// backward of it) is run layer by layer as 128-sample x N_out GEMMs on the 5th-generation tensor
// cores: `tcgen05.mma.cta_group::1.kind::tf32`, operands read from shared memory through matrix
// descriptors, fp32 accumulators in tensor memory, one elected thread issuing.  fp32 parity comes
// from the same error-compensated split the mma.sync kernel uses (x = hi + lo, hi = x with the low
// 13 mantissa bits cleared): D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
//
//   umma_weight_images_kernel   W_k -> swizzled (hi, lo) operand images, forward and transposed
//   umma_gemm_kernel<EPI>       D[128 x N] = A[128 x K] * Bimg[N x K]^T, warp-specialised:
//                               4 producer warps (global/gather -> registers -> swizzled smem, split),
//                               1 MMA warp, 4 epilogue warps (TMEM -> registers -> fused epilogue)
//       EPI_RELU_STORE  forward layer:  act[k+1] = relu(D + b)
//       EPI_PREDICT     last layer:     h_L = relu(D + b); logit, loss, dlogit, predict grads,
//                                       GMF-branch scatter, delta_L
//       EPI_MASK_STORE  backward data:  delta[k] = D * (act[k] > 0)
//       EPI_SCATTER     backward data of layer 0: D -> RED.128 into the embedding-gradient rows
//   umma_wgrad_kernel           dW_k = delta[k+1]^T act[k] summed over the CTA's samples with the
//                               accumulator resident in TMEM; both operands MN-major (128B_BASE32B)
//
// Layouts verified on hardware by tools/umma_probe.cu (profiles/umma_probe_r01.txt).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tile_params.cuh"

namespace {

constexpr int kTile = 128;  // samples per accumulator tile (UMMA M)
constexpr uint32_t kHiMask = 0xffffe000u;
constexpr int kProducerWarps = 8;
constexpr int kDepth = 3;           // panels of loads in flight per producer thread
constexpr int kGemmThreads = 544;   // warps 0-7 producers, 8 MMA, 9-12 / 13-16 epilogue (even / odd tiles)
constexpr int kChunk = 32;          // samples per wgrad stage (4 k-steps of 8)
constexpr int kMaxStages = 4;

enum { EPI_RELU_STORE = 0, EPI_PREDICT = 1, EPI_PREDICT_TRAIN = 2, EPI_MASK_STORE = 3, EPI_SCATTER = 4 };

struct GemmArgs {
  int K;        // reduction width (multiple of 8)
  int N;        // output columns of one block (<= 256, multiple of 16)
  int n_total;  // rows of the weight image / leading dimension of the output
  int passes;   // 3 = fp32-parity split, 1 = plain TF32
  int layer;    // tower layer index k
  int gather;   // A rows are [embed_user_MLP[u] | embed_item_MLP[i]] instead of a dense matrix
  int stages;
  int ablate;   // debugging knock-outs (NCF_UMMA_ABLATE bit mask), 0 in production
  const float* a;      // dense A [B][K]
  const float* b_img;  // weight image: [panel][hi|lo][n_total rows][32]
  const float* bias;
  float* out;
  const float* mask_src;
};

struct WgradJob {
  int k, mb, nb, cta0, nctas;
  int S;  // samples per stage: scaled so that every job moves the same bytes per hand-off
};
struct WgradArgs {
  int njobs, passes, ablate;
  WgradJob job[12];
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (launch failure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// One lane polls / arrives for its warp: 32 arrivals on one mbarrier serialise in shared memory.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* b, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(b, parity);
  __syncwarp();
}
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* b) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(b);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 16-byte global -> shared copy through L2 only; !valid writes zeros (src-size 0, nothing is read)
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive columns: thread = TMEM lane (row), r[j] = column j
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// shared-memory matrix descriptor (sm_100 version 1): start, leading/stride byte offsets, layout type
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)(layout & 7) << 61;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: kind::tf32, fp32 accumulate, M x N, operand majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (uint32_t)a_mn << 15;
  d |= (uint32_t)b_mn << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(x.x) & kHiMask);
  hi.y = __uint_as_float(__float_as_uint(x.y) & kHiMask);
  hi.z = __uint_as_float(__float_as_uint(x.z) & kHiMask);
  hi.w = __uint_as_float(__float_as_uint(x.w) & kHiMask);
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// Column sums over the warp: v[j] = this lane's value of column j; lane c returns the sum of column c.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? v[i] : v[i + o];
      const float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- weight images -----------------------------------------------------------------------------------
// K-major SWIZZLE_128B image of a logical [rows][K] matrix: panel pi = 32 columns,
//   [pi][hi|lo][row][128 B], 16-byte chunk index ^= row % 8.
// dir 0: rows = W[k+1] (n_out), K = W[k]   (forward B operand:  elem(n, c) = w[n][c])
// dir 1: rows = W[k]   (k_in),  K = W[k+1] (backward-data B operand: elem(n, c) = w[c][n])
__global__ void umma_weight_images_kernel(const TileParams p) {
  const int k = blockIdx.y >> 1, dir = blockIdx.y & 1;
  const int kin = p.W[k], nout = p.W[k + 1];
  const int rows = dir ? kin : nout, K = dir ? nout : kin;
  const int panels = (K + 31) / 32;
  const float* w = p.w[k];
  float* img = dir ? p.wsplit_b[k] : p.wsplit_f[k];
  const int64_t total = (int64_t)panels * rows * 32;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx & 31);
    const int r = (int)((idx >> 5) % rows);
    const int pi = (int)((idx >> 5) / rows);
    const int c = pi * 32 + cc;
    float x = 0.f;
    if (c < K) x = dir ? w[(int64_t)c * kin + r] : w[(int64_t)r * kin + c];
    const float hi = __uint_as_float(__float_as_uint(x) & kHiMask);
    const int64_t off = (int64_t)pi * 2 * rows * 32 + (int64_t)r * 32 + ((((cc >> 2) ^ (r & 7)) << 2) | (cc & 3));
    img[off] = hi;
    img[off + (int64_t)rows * 32] = x - hi;
  }
}

#ifdef NCF_UMMA_TRACE
// timeline of CTA 0 (debug builds only): clock64 at the hand-off points of each role
__device__ long long g_trace[3][128];
#define NCF_TRACE(role, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 128) g_trace[role][slot] = clock64(); } while (0)
#else
#define NCF_TRACE(role, slot) do { } while (0)
#endif

// ---- GEMM kernel ------------------------------------------------------------------------------------
struct GemmBars {
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  float pg[2 * 128 + 4];  // predict-weight / predict-bias gradient of this CTA (EPI_PREDICT_TRAIN)
};

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ TileParams p, const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ GemmBars bars;
  // 1024-byte alignment of the swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = g.K, N = g.N;
  const int panels = (K + 31) >> 5;
  const uint32_t a_bytes = kTile * 128, b_bytes = (uint32_t)N * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const int nb = blockIdx.y;  // output-column block
  const int64_t ntiles = (p.B + kTile - 1) / kTile;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * N) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < g.stages; ++s) { mbar_init(&bars.full[s], kProducerWarps + 1); mbar_init(&bars.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars.acc_full[a], 1); mbar_init(&bars.acc_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (EPI == EPI_PREDICT_TRAIN)
    for (int i = tid; i < 2 * 128 + 4; i += kGemmThreads) bars.pg[i] = 0.f;
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp < kProducerWarps) {
    // ===== producers: two threads per sample row, each moves 64 B of every 128-byte panel row; the
    // (tile, panel) sequence is one flat software pipeline with kDepth panels of loads in flight ===========
    const int t = tid, r = t >> 1, h = t & 1;
    const int d = p.d;
    const int64_t my_tiles = (blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_iter = my_tiles * panels;
    // load-side cursor
    int64_t ld_tl = 0;
    int ld_pi = 0;
    bool ld_ok = false;
    const float *ld_u = nullptr, *ld_i = nullptr;
    int64_t pre_u = -1, pre_it = -1;  // indices of the load cursor's NEXT tile, fetched one tile early
    auto fetch_idx = [&](int64_t tl) {
      pre_u = -1;
      pre_it = -1;
      if (tl < my_tiles) {
        const int64_t row = (blockIdx.x + tl * gridDim.x) * kTile + r;
        if (row < p.B) {
          pre_u = p.user[p.user_div > 0 ? row / p.user_div : row];
          pre_it = p.item[row];
        }
      }
    };
    auto enter_tile = [&](int64_t tl) {
      const int64_t row = (blockIdx.x + tl * gridDim.x) * kTile + r;
      ld_ok = row < p.B;
      if (g.gather) {
        const int64_t u = pre_u, it = pre_it;
        if (u < 0 || u >= p.U || it < 0 || it >= p.I) ld_ok = false;
        else { ld_u = p.eum + u * d; ld_i = p.eim + it * d; }
        fetch_idx(tl + 1);
      } else if (ld_ok) {
        ld_u = g.a + row * (int64_t)K;
      }
    };
    auto issue = [&](float4 (&v)[4]) {
      if (ld_pi == 0) enter_tile(ld_tl);
      const int c0 = ld_pi * 32 + h * 16;
      const float* src = nullptr;
      if (ld_ok && !(g.ablate & 1)) src = g.gather ? (c0 < d ? ld_u + c0 : ld_i + (c0 - d)) : ld_u + c0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (src != nullptr && c0 + c * 4 < K) v[c] = ldg4(src + c * 4);
        else v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++ld_pi == panels) { ld_pi = 0; ++ld_tl; }
    };
    int st_pi = 0, s = 0;
    uint32_t ph = 0;
    int tr = 0;
    auto consume = [&](const float4 (&v)[4]) {
      mbar_wait_warp(&bars.empty[s], ph ^ 1);
      if (t == 0) NCF_TRACE(0, tr);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      if (t == 0) {
        const float* bsrc = g.b_img + (int64_t)st_pi * 2 * g.n_total * 32 + (int64_t)nb * N * 32;
        if (g.ablate & 2) {
          mbar_arrive(&bars.full[s]);
        } else if (g.passes == 3) {
          mbar_expect_tx(&bars.full[s], 2 * b_bytes);
          bulk_g2s(st + 2 * a_bytes, bsrc, b_bytes, &bars.full[s]);
          bulk_g2s(st + 2 * a_bytes + b_bytes, bsrc + (int64_t)g.n_total * 32, b_bytes, &bars.full[s]);
        } else {
          mbar_expect_tx(&bars.full[s], b_bytes);
          bulk_g2s(st + 2 * a_bytes, bsrc, b_bytes, &bars.full[s]);
        }
      }
      uint8_t* a_hi = st + r * 128;
      uint8_t* a_lo = a_hi + a_bytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 hi, lo;
        split4(v[c], hi, lo);
        const int off = ((h * 4 + c) ^ (r & 7)) << 4;
        if (g.ablate & 16) continue;
        *reinterpret_cast<float4*>(a_hi + off) = v[c];  // the tensor core ignores the low 13 mantissa bits
        if (g.passes == 3) *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
      fence_async_smem();
      if (t == 0) NCF_TRACE(0, tr + 1);
      tr += 2;
      mbar_arrive_warp(&bars.full[s]);
      if (++st_pi == panels) st_pi = 0;
      if (++s == g.stages) { s = 0; ph ^= 1; }
    };
    if (g.gather) fetch_idx(0);
    float4 buf[kDepth][4];
#pragma unroll
    for (int j = 0; j < kDepth; ++j)
      if (j < n_iter) issue(buf[j]);
    for (int64_t i0 = 0; i0 < n_iter; i0 += kDepth) {
#pragma unroll
      for (int j = 0; j < kDepth; ++j) {
        const int64_t i = i0 + j;
        if (i < n_iter) {
          consume(buf[j]);
          if (i + kDepth < n_iter) issue(buf[j]);
        }
      }
    }
  } else if (warp == kProducerWarps) {
    // ===== MMA issuer =======================================================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kTile, N, 0, 0);
      uint32_t lt = 0, ph = 0;
      int s = 0, tr = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
        const int a = lt & 1;
        mbar_wait(&bars.acc_empty[a], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + a * N;
        for (int pi = 0; pi < panels; ++pi) {
          mbar_wait(&bars.full[s], ph);
          NCF_TRACE(1, tr);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t ahi = base, alo = base + a_bytes, bhi = base + 2 * a_bytes, blo = bhi + b_bytes;
          const int ksteps = (g.ablate & 4) ? 0 : min(4, (K - pi * 32) >> 3);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dah = make_desc(ahi + ks * 32, 16, 1024, 2);
            const uint64_t dbh = make_desc(bhi + ks * 32, 16, 1024, 2);
            tc_mma(d_tmem, dah, dbh, idesc, (pi > 0 || ks > 0) ? 1u : 0u);
            if (g.passes == 3) {
              const uint64_t dal = make_desc(alo + ks * 32, 16, 1024, 2);
              const uint64_t dbl = make_desc(blo + ks * 32, 16, 1024, 2);
              tc_mma(d_tmem, dal, dbh, idesc, 1u);
              tc_mma(d_tmem, dah, dbl, idesc, 1u);
            }
          }
          tc_commit(&bars.empty[s]);
          NCF_TRACE(1, tr + 1);
          tr += 2;
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
        tc_commit(&bars.acc_full[a]);
      }
    }
  } else {
    // ===== epilogue: warp quarter q owns TMEM lanes [32q, 32q+32) ============================================
    const int q = warp & 3;
    const int grp = (warp - kProducerWarps - 1) >> 2;  // epilogue group: even / odd local tiles
    const int f = p.f, dmlp = p.d;
    const int mlp_off = (p.type == NCF_NEUMF) ? f : 0;
    const bool has_gmf = p.type != NCF_MLP;
    uint32_t lt = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int a = lt & 1;
      if (a != grp) continue;
      mbar_wait_warp(&bars.acc_full[a], (lt >> 1) & 1);
      if (q == 0 && lane == 0) NCF_TRACE(2, 2 * (int)lt);
      tc_fence_after();
      const int64_t row = tile * kTile + q * 32 + lane;
      const bool valid = row < p.B;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + a * N;
      float v[32];
      if (g.ablate & 8) {
        tc_fence_before();
        mbar_arrive_warp(&bars.acc_empty[a]);
        continue;
      }

      if (EPI == EPI_RELU_STORE) {
        float* out = g.out + row * (int64_t)g.n_total + nb * N;
        const float* bias = g.bias + nb * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
          tc_ld32(taddr + c0, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = ldg4(bias + c0 + j);
              *reinterpret_cast<float4*>(out + c0 + j) =
                  make_float4(fmaxf(v[j] + bb.x, 0.f), fmaxf(v[j + 1] + bb.y, 0.f),
                              fmaxf(v[j + 2] + bb.z, 0.f), fmaxf(v[j + 3] + bb.w, 0.f));
            }
          }
        }
      } else if (EPI == EPI_MASK_STORE) {
        float* out = g.out + row * (int64_t)g.n_total + nb * N;
        const float* h = g.mask_src + row * (int64_t)g.n_total + nb * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
          tc_ld32(taddr + c0, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 hh = ldg4(h + c0 + j);
              *reinterpret_cast<float4*>(out + c0 + j) =
                  make_float4(hh.x > 0.f ? v[j] : 0.f, hh.y > 0.f ? v[j + 1] : 0.f,
                              hh.z > 0.f ? v[j + 2] : 0.f, hh.w > 0.f ? v[j + 3] : 0.f);
            }
          }
        }
      } else if (EPI == EPI_SCATTER) {
        int64_t u = -1, it = -1;
        if (valid) {
          u = p.user[p.user_div > 0 ? row / p.user_div : row];
          it = p.item[row];
          if (u < 0 || u >= p.U || it < 0 || it >= p.I) u = -1;
        }
        for (int c0 = 0; c0 < N; c0 += 32) {
          tc_ld32(taddr + c0, v);
          if (u >= 0) {
            const int col = nb * N + c0;
            float* dst = (col < dmlp) ? p.gum + u * dmlp + col : p.gim + it * dmlp + (col - dmlp);
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      } else {  // EPI_PREDICT / EPI_PREDICT_TRAIN: N == f (the last tower layer), nb == 0
        constexpr bool TRAIN = (EPI == EPI_PREDICT_TRAIN);
        int64_t u = -1, it = -1;
        bool bad = false;
        if (valid) {
          u = p.user[p.user_div > 0 ? row / p.user_div : row];
          it = p.item[row];
          if (u < 0 || u >= p.U || it < 0 || it >= p.I) { u = -1; bad = true; }
        }
        const bool ok = u >= 0;
        float acc = 0.f;
        for (int c0 = 0; c0 < N; c0 += 32) {
          tc_ld32(taddr + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc = fmaf(__ldg(&p.pw[mlp_off + c0 + j]), fmaxf(v[j] + __ldg(&g.bias[c0 + j]), 0.f), acc);
        }
        if (has_gmf && ok) {
          const float* ru = p.eug + u * f;
          const float* ri = p.eig + it * f;
          for (int c = 0; c < f; c += 4) {
            const float4 gu = ldg4(ru + c), gi = ldg4(ri + c), w = ldg4(p.pw + c);
            acc = fmaf(w.x, gu.x * gi.x, acc);
            acc = fmaf(w.y, gu.y * gi.y, acc);
            acc = fmaf(w.z, gu.z * gi.z, acc);
            acc = fmaf(w.w, gu.w * gi.w, acc);
          }
        }
        float x = acc + __ldg(p.pb);
        if (bad) x = __int_as_float(0x7fc00000);  // out-of-range index: NaN
        if (valid && p.logits != nullptr) p.logits[row] = x;
        if (TRAIN) {
          float dl = 0.f, ls = 0.f;
          if (ok && p.dlogit_in != nullptr) {
            dl = p.dlogit_in[row];
          } else if (ok) {
            const float y = p.label[row];
            const float e = expf(-fabsf(x));
            const float bce = fmaxf(x, 0.f) - x * y + log1pf(e);
            const float sig = (x >= 0.f) ? 1.f / (1.f + e) : e / (1.f + e);
            if (p.teacher != nullptr) {
              const float df = x - p.teacher[row];
              ls = p.alpha * bce + (1.f - p.alpha) * df * df;
              dl = (p.alpha * (sig - y) + (1.f - p.alpha) * 2.f * df) * p.invB;
            } else {
              ls = bce;
              dl = (sig - y) * p.invB;
            }
          }
          const float ls_w = warp_sum(ls), dl_w = warp_sum(dl);
          if (lane == 0) {
            if (p.loss_accum != nullptr) atomicAdd(p.loss_accum, (double)ls_w * (double)p.invB);
            atomicAdd(&bars.pg[p.predict_size], dl_w);
          }
          // delta_L and the predict-weight gradient of the tower half
          float* dz = g.out + row * (int64_t)N;
          for (int c0 = 0; c0 < N; c0 += 32) {
            tc_ld32(taddr + c0, v);
            float z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float h = fmaxf(v[j] + __ldg(&g.bias[c0 + j]), 0.f);
              z[j] = (h > 0.f) ? dl * __ldg(&p.pw[mlp_off + c0 + j]) : 0.f;
              v[j] = dl * h;
            }
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dz + c0 + j) = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
            }
            const float s = warp_colsum32(v, lane);
            atomicAdd(&bars.pg[mlp_off + c0 + lane], s);
          }
          // GMF branch: predict-weight gradient and the scatter into the GMF embedding-gradient rows
          if (has_gmf) {
            for (int c0 = 0; c0 < f; c0 += 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 gu = make_float4(0.f, 0.f, 0.f, 0.f), gi = gu;
                if (ok) {
                  gu = ldg4(p.eug + u * f + c0 + j);
                  gi = ldg4(p.eig + it * f + c0 + j);
                  const float4 w = ldg4(p.pw + c0 + j);
                  const float4 wd = make_float4(w.x * dl, w.y * dl, w.z * dl, w.w * dl);
                  red_add4(p.gug + u * f + c0 + j, make_float4(wd.x * gi.x, wd.y * gi.y, wd.z * gi.z, wd.w * gi.w));
                  red_add4(p.gig + it * f + c0 + j, make_float4(wd.x * gu.x, wd.y * gu.y, wd.z * gu.z, wd.w * gu.w));
                }
                v[j] = dl * (gu.x * gi.x);
                v[j + 1] = dl * (gu.y * gi.y);
                v[j + 2] = dl * (gu.z * gi.z);
                v[j + 3] = dl * (gu.w * gi.w);
              }
              const float s = warp_colsum32(v, lane);
              atomicAdd(&bars.pg[c0 + lane], s);
            }
          }
        }
      }
      tc_fence_before();
      if (q == 0 && lane == 0) NCF_TRACE(2, 2 * (int)lt + 1);
      mbar_arrive_warp(&bars.acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_PREDICT_TRAIN)
    for (int i = tid; i <= p.predict_size; i += kGemmThreads) atomicAdd(&p.gt[p.pw_off + i], bars.pg[i]);
  if (warp == kProducerWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

// ---- weight-gradient kernel -----------------------------------------------------------------------------
// dW_k[n][c] = sum_s delta[k+1][s][n] * act[k][s][c].  The reduction runs over samples, the slow
// dimension of both operands, so both are MN-major: image rows = samples (128 B = 32 features per
// row and panel), 4-row atoms with the 32-byte chunk index ^= row % 4 (128B_BASE32B, the only
// MN-major layout the tf32 kind accepts).  One CTA owns one [<=128 x <=256] block of one layer's dW
// for its share of the batch; the accumulator stays in TMEM until the CTA's last chunk.
constexpr int kWgProducerWarps = 8;
constexpr int kWgProducers = kWgProducerWarps * 32;
constexpr int kWgradThreads = kWgProducers + 32;  // + the MMA warp

struct WgradBars {
  uint64_t full[2], empty[2], acc_full;
  uint32_t tmem_base;
  float db[128];
};

// byte offset of feature fo (multiple of 4) of sample row s in an image of S rows per panel
__device__ __forceinline__ uint32_t mn_offset(int s, int fo, int S) {
  const int panel = fo >> 5, cc = fo & 31;
  return (uint32_t)(panel * (S * 128) + s * 128 + (((cc >> 3) ^ (s & 3)) << 5) + ((cc & 7) << 2));
}

__global__ void __launch_bounds__(kWgradThreads, 1)
umma_wgrad_kernel(const __grid_constant__ TileParams p, const __grid_constant__ WgradArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ WgradBars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ji = 0;
  while (ji + 1 < g.njobs && (int)blockIdx.x >= g.job[ji + 1].cta0) ++ji;
  const WgradJob job = g.job[ji];
  const int k = job.k, kin = p.W[k], nout = p.W[k + 1];
  const int MB = min(128, nout - job.mb * 128), NB = min(256, kin - job.nb * 256);
  const int local = (int)blockIdx.x - job.cta0;
  const int S = job.S;
  const int64_t nchunks = (p.B + S - 1) / S;
  const uint32_t a_img = (uint32_t)MB * S * 4, b_img = (uint32_t)NB * S * 4;
  const uint32_t stage_bytes = 2 * a_img + 2 * b_img;
  const int64_t my_chunks = (local < nchunks) ? (nchunks - local + job.nctas - 1) / job.nctas : 0;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&bars.full[s], kWgProducerWarps); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) bars.db[tid] = 0.f;
  if (warp == kWgProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&bars.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp < kWgProducerWarps) {
    // Producers: every thread copies its 16-byte pieces of the chunk global -> shared with cp.async
    // (raw fp32 = the "hi" image: the tensor core ignores the low 13 mantissa bits), one chunk ahead
    // of the chunk whose "lo" image it is computing from the pieces it copied itself.
    const int t = tid, d = p.d;
    const float* delta = p.delta[k + 1];
    const float* actk = (k > 0) ? p.act[k] : nullptr;
    const int pa = MB >> 2, pb = NB >> 2;       // 16-byte pieces per sample row
    const int na = S * pa / kWgProducers, nbp = S * pb / kWgProducers;  // pieces per thread and chunk (<= 4, <= 8)
    const int ca = t % pa, sa0 = t / pa, sas = kWgProducers / pa;
    const int cb = t % pb, sb0 = t / pb, sbs = kWgProducers / pb;
    float4 dbsum = make_float4(0.f, 0.f, 0.f, 0.f);
    // layer 0: the activation operand is gathered.  A thread always reads the same 4 columns, so it
    // needs one of the two tables only; its sample indices are fetched one chunk ahead so that the
    // row copies never wait on an index load.
    const int fo = job.nb * 256 + cb * 4;
    const bool from_user = fo < d;
    const int64_t* idx_src = from_user ? p.user : p.item;
    const float* tab = from_user ? p.eum + fo : p.eim + (fo - d);
    const int64_t idx_lim = from_user ? p.U : p.I;
    int64_t idxv[8];
    auto issue_idx = [&](int64_t ci) {
      const int64_t row0 = (local + ci * job.nctas) * S;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = row0 + sb0 + i * sbs;
        idxv[i] = (i < nbp && ci < my_chunks && row < p.B) ? idx_src[row] : -1;
      }
    };
    auto issue = [&](int64_t ci) {
      const int s = (int)(ci & 1);
      mbar_wait_warp(&bars.empty[s], (uint32_t)((ci >> 1) & 1) ^ 1);
      if (t == 0) NCF_TRACE(0, 2 * (int)ci);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      const int64_t row0 = (local + ci * job.nctas) * S;
      const bool off = (g.ablate & 64) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < na) {
          const int64_t row = row0 + sa0 + i * sas;
          const bool ok = row < p.B && !off;
          cp_async16_zfill(st + mn_offset(sa0 + i * sas, ca * 4, S),
                           ok ? delta + row * nout + job.mb * 128 + ca * 4 : delta, ok);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < nbp) {
          const float* src;
          bool ok;
          if (k == 0) {
            ok = idxv[i] >= 0 && idxv[i] < idx_lim && !off;
            src = ok ? tab + idxv[i] * d : tab;
          } else {
            const int64_t row = row0 + sb0 + i * sbs;
            ok = row < p.B && !off;
            src = ok ? actk + row * kin + fo : actk;
          }
          cp_async16_zfill(st + 2 * a_img + mn_offset(sb0 + i * sbs, cb * 4, S), src, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (k == 0) issue_idx(ci + 1);
    };
    auto consume = [&](int64_t ci, bool more_pending) {
      const int s = (int)(ci & 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      if (more_pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (!(g.ablate & 256)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < na) {
            const uint32_t off = mn_offset(sa0 + i * sas, ca * 4, S);
            const float4 x = *reinterpret_cast<const float4*>(st + off);
            float4 hi, lo;
            split4(x, hi, lo);
            dbsum.x += x.x; dbsum.y += x.y; dbsum.z += x.z; dbsum.w += x.w;
            if (g.ablate & 512) *reinterpret_cast<float4*>(st + off) = x;
            if (g.passes == 3) *reinterpret_cast<float4*>(st + a_img + off) = lo;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nbp && g.passes == 3) {
            const uint32_t off = 2 * a_img + mn_offset(sb0 + i * sbs, cb * 4, S);
            const float4 x = *reinterpret_cast<const float4*>(st + off);
            float4 hi, lo;
            split4(x, hi, lo);
            if (g.ablate & 512) *reinterpret_cast<float4*>(st + off) = x;
            *reinterpret_cast<float4*>(st + b_img + off) = lo;
          }
        }
      }
      fence_async_smem();
      if (t == 0) NCF_TRACE(0, 2 * (int)ci + 1);
      mbar_arrive_warp(&bars.full[s]);
    };
    if (k == 0) issue_idx(0);
    if (my_chunks > 0) issue(0);
    for (int64_t ci = 0; ci < my_chunks; ++ci) {
      const bool more = ci + 1 < my_chunks;
      if (more) issue(ci + 1);
      consume(ci, more);
    }
    // bias gradient: column sums of delta (thread t always sees the same 4 columns)
    if (job.nb == 0) {
      atomicAdd(&bars.db[ca * 4 + 0], dbsum.x);
      atomicAdd(&bars.db[ca * 4 + 1], dbsum.y);
      atomicAdd(&bars.db[ca * 4 + 2], dbsum.z);
      atomicAdd(&bars.db[ca * 4 + 3], dbsum.w);
    }
    named_bar(1, kWgProducers);
    if (job.nb == 0 && t < MB && my_chunks > 0) atomicAdd(&p.gt[p.b_off[k] + job.mb * 128 + t], bars.db[t]);
    // flush the accumulator block: warp w reads TMEM lanes 32*(w%4).., column half w/4
    if (my_chunks > 0) {
      mbar_wait_warp(&bars.acc_full, 0);
      tc_fence_after();
      const int q = warp & 3, part = warp >> 2, parts = min(kWgProducerWarps / 4, NB / 32);
      const int m = q * 32 + lane;
      float* dst = p.gt + p.w_off[k] + (int64_t)(job.mb * 128 + m) * kin + job.nb * 256;
      float v[32];
      for (int c0 = part * (NB / parts); c0 < (part + 1) * (NB / parts) && part < parts; c0 += 32) {
        tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
        if (m < MB && !(g.ablate & 32)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) red_add4(dst + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = make_idesc(128, NB, 1, 1);
    for (int64_t ci = 0; ci < my_chunks; ++ci) {
      const int s = (int)(ci & 1);
      mbar_wait(&bars.full[s], (uint32_t)(ci >> 1) & 1);
      NCF_TRACE(1, 2 * (int)ci);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t ahi = base, alo = base + a_img, bhi = base + 2 * a_img, blo = bhi + b_img;
      for (int ks = 0; ks < ((g.ablate & 128) ? 0 : S / 8); ++ks) {
        const uint64_t dah = make_desc(ahi + ks * 1024, S * 128, 512, 1);
        const uint64_t dbh = make_desc(bhi + ks * 1024, S * 128, 512, 1);
        tc_mma(tmem, dah, dbh, idesc, (ci > 0 || ks > 0) ? 1u : 0u);
        if (g.passes == 3) {
          const uint64_t dal = make_desc(alo + ks * 1024, S * 128, 512, 1);
          const uint64_t dbl = make_desc(blo + ks * 1024, S * 128, 512, 1);
          tc_mma(tmem, dal, dbh, idesc, 1u);
          tc_mma(tmem, dah, dbl, idesc, 1u);
        }
      }
      tc_commit(&bars.empty[s]);
      NCF_TRACE(1, 2 * (int)ci + 1);
    }
    if (my_chunks > 0) tc_commit(&bars.acc_full);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgProducerWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ---- launchers ---------------------------------------------------------------------------------------------
constexpr size_t kSmemBudget = 200 * 1024;

template <int EPI>
int launch_gemm(const TileParams& p, GemmArgs g, int nblocks, cudaStream_t st) {
  const size_t stage = 2 * (size_t)kTile * 128 + 2 * (size_t)g.N * 128;
  int stages = (int)(kSmemBudget / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    ncf::set_error("umma gemm: stage of %zu bytes does not fit twice", stage);
    return NCF_ERR_ARG;
  }
  g.stages = stages;
  if (const char* ab = getenv("NCF_UMMA_ABLATE")) g.ablate = atoi(ab);
  const size_t smem = stages * stage + 1024;
  auto kern = umma_gemm_kernel<EPI>;
  NCF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (p.B + kTile - 1) / kTile;
  int64_t grid = ncf::num_sms();
  if (grid > ntiles) grid = ntiles;
  kern<<<dim3((unsigned)grid, (unsigned)nblocks), kGemmThreads, smem, st>>>(p, g);
  NCF_LAUNCH_CHECK("umma_gemm_kernel");
#ifdef NCF_UMMA_TRACE
  {
    static int calls = 0;
    const char* want = getenv("NCF_UMMA_TRACE_CALL");  // index of the launch to print (counted per process)
    if (want && atoi(want) == calls) {
      cudaStreamSynchronize(st);
      static long long h[3][128];
      cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
      const long long t0 = h[0][0];
      fprintf(stderr, "[trace] EPI=%d K=%d N=%d stages=%d\n", EPI, g.K, g.N, g.stages);
      for (int i = 0; i < 48; i += 2)
        fprintf(stderr, "[trace] %2d  prod: empty-ok %7lld stored %7lld | mma: full-ok %7lld committed %7lld | epi(tile %d): acc-ok %7lld done %7lld\n",
                i / 2, h[0][i] - t0, h[0][i + 1] - t0, h[1][i] - t0, h[1][i + 1] - t0, i / 2,
                i / 2 < 8 ? h[2][i] - t0 : 0, i / 2 < 8 ? h[2][i + 1] - t0 : 0);
    }
    ++calls;
  }
#endif
  return NCF_OK;
}

int launch_wgrad(const TileParams& p, int passes, cudaStream_t st) {
  WgradArgs g{};
  g.passes = passes;
  if (const char* ab = getenv("NCF_UMMA_ABLATE")) g.ablate = atoi(ab);
  // one job per [128 x 256] block of every layer's dW; CTAs shared out by operand bytes per sample
  double weight[12], total = 0;
  for (int k = 0; k < p.L; ++k) {
    const int mbs = (p.W[k + 1] + 127) / 128, nbs = (p.W[k] + 255) / 256;
    for (int mb = 0; mb < mbs; ++mb)
      for (int nb = 0; nb < nbs; ++nb) {
        if (g.njobs >= 12) {
          ncf::set_error("umma wgrad: too many weight blocks");
          return NCF_ERR_ARG;
        }
        const int MB = std::min(128, p.W[k + 1] - mb * 128), NB = std::min(256, p.W[k] - nb * 256);
        int S = std::min(4096 / MB, 8192 / NB) / 8 * 8;  // <= 4 + 8 sixteen-byte pieces per producer thread
        S = std::max(8, std::min(S, 256));
        g.job[g.njobs] = WgradJob{k, mb, nb, 0, 0, S};
        weight[g.njobs] = 1.0 / S;  // hand-offs per sample; every hand-off costs about the same
        total += weight[g.njobs];
        ++g.njobs;
      }
  }
  const int sms = ncf::num_sms();
  int cta = 0;
  size_t smem = 0;
  for (int j = 0; j < g.njobs; ++j) {
    const int64_t nchunks = (p.B + g.job[j].S - 1) / g.job[j].S;
    int n = (int)(weight[j] / total * sms);
    if (n < 1) n = 1;
    if (n > nchunks) n = (int)nchunks;
    g.job[j].cta0 = cta;
    g.job[j].nctas = n;
    cta += n;
    const int k = g.job[j].k;
    const int MB = std::min(128, p.W[k + 1] - g.job[j].mb * 128), NB = std::min(256, p.W[k] - g.job[j].nb * 256);
    smem = std::max(smem, (size_t)2 * 2 * (MB + NB) * g.job[j].S * 4);
  }
  smem += 1024 + 16 * 1024;  // alignment slack + the M=128 descriptor may read past a narrow A' image
  NCF_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_wgrad_kernel<<<cta, kWgradThreads, smem, st>>>(p, g);
  NCF_LAUNCH_CHECK("umma_wgrad_kernel");
#ifdef NCF_UMMA_TRACE
  if (getenv("NCF_UMMA_TRACE_WGRAD")) {
    cudaStreamSynchronize(st);
    static long long h[3][128];
    cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
    const long long t0 = h[0][0];
    for (int i = 0; i < 48; i += 2)
      fprintf(stderr, "[wtrace] %2d  prod: empty-ok %7lld stored %7lld | mma: full-ok %7lld committed %7lld\n", i / 2,
              h[0][i] - t0, h[0][i + 1] - t0, h[1][i] - t0, h[1][i + 1] - t0);
  }
#endif
  return NCF_OK;
}

// NCF_UMMA_TIMING=1: CUDA-event time of every launch of one training step, printed to stderr
// (debugging aid; synchronises the stream).
struct StepTimer {
  bool on;
  cudaStream_t st;
  cudaEvent_t ev[32];
  const char* name[32];
  int n = 0;
  static thread_local StepTimer* g_timer;
  StepTimer(cudaStream_t s) : on(getenv("NCF_UMMA_TIMING") != nullptr), st(s) {
    g_timer = this;
    mark("start");
  }
  void mark(const char* what) {
    if (!on || n >= 32) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  ~StepTimer() {
    g_timer = nullptr;
    if (!on) return;
    cudaStreamSynchronize(st);
    for (int i = 1; i < n; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, "[umma] %-10s %8.1f us\n", name[i], ms * 1e3f);
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

thread_local StepTimer* StepTimer::g_timer = nullptr;

GemmArgs forward_args(const TileParams& p, int k, int passes) {
  GemmArgs g{};
  g.K = p.W[k];
  g.n_total = p.W[k + 1];
  g.N = std::min(256, g.n_total);
  g.passes = passes;
  g.layer = k;
  g.gather = (k == 0);
  g.a = (k == 0) ? nullptr : p.act[k];
  g.b_img = p.wsplit_f[k];
  g.bias = p.b[k];
  return g;
}

int tower_forward(TileParams& p, int passes, bool train, cudaStream_t st) {
  for (int k = 0; k < p.L; ++k) {
    GemmArgs g = forward_args(p, k, passes);
    int rc;
    if (k + 1 < p.L) {
      g.out = p.act[k + 1];
      rc = launch_gemm<EPI_RELU_STORE>(p, g, g.n_total / g.N, st);
    } else if (train) {
      g.out = p.delta[p.L];
      rc = launch_gemm<EPI_PREDICT_TRAIN>(p, g, 1, st);
    } else {
      rc = launch_gemm<EPI_PREDICT>(p, g, 1, st);
    }
    if (rc != NCF_OK) return rc;
    if (StepTimer::g_timer) StepTimer::g_timer->mark(k == 0 ? "fwd0" : k == 1 ? "fwd1" : "fwd2+");
  }
  return NCF_OK;
}

}  // namespace

namespace ncf {

// Eligible: a tower whose widths are power-of-two multiples of 32 (panels of 32 fp32, accumulator
// chunks of 32 columns) and a batch large enough that ~10 launches beat one fused launch.
bool umma_eligible(const TileParams& p) {
  const char* off = getenv("NCF_UMMA_DISABLE");
  if (off != nullptr && off[0] == '1') return false;
  const char* mb = getenv("NCF_UMMA_MIN_B");
  const int64_t min_b = mb ? atoll(mb) : 8192;
  if (p.B < min_b) return false;
  if (p.type == NCF_GMF) return false;
  if (p.f < 32 || p.f > 128 || (p.f & (p.f - 1)) != 0) return false;
  return true;
}

int64_t umma_image_floats(const TileParams& p) {
  int64_t n = 0;
  for (int k = 0; k < p.L; ++k) n += 2 * 2 * (int64_t)p.W[k] * p.W[k + 1];
  return n;
}

int64_t umma_scratch_floats(const TileParams& p, int64_t B, bool train) {
  int64_t n = 0;
  for (int k = 1; k < p.L; ++k) n += B * p.W[k];
  if (train)
    for (int k = 1; k <= p.L; ++k) n += B * p.W[k];
  return n;
}

constexpr int64_t kForwardSub = 262144;  // inference sub-batch: bounds the activation scratch

int64_t umma_forward_workspace_floats(const TileParams& p, int64_t B) {
  return umma_image_floats(p) + umma_scratch_floats(p, std::min<int64_t>(B, kForwardSub), false) + 64;
}
int64_t umma_train_workspace_floats(const TileParams& p, int64_t B) {
  return umma_image_floats(p) + umma_scratch_floats(p, B, true) + 64;
}

static float* carve(TileParams& p, float* ws, int64_t B, bool train, cudaStream_t st, int* rc) {
  for (int k = 0; k < p.L; ++k) {
    const int64_t n = 2 * (int64_t)p.W[k] * p.W[k + 1];
    p.wsplit_f[k] = ws;
    p.wsplit_b[k] = ws + n;
    ws += 2 * n;
  }
  ws = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  for (int k = 1; k < p.L; ++k) { p.act[k] = ws; ws += B * p.W[k]; }
  if (train)
    for (int k = 1; k <= p.L; ++k) { p.delta[k] = ws; ws += B * p.W[k]; }
  umma_weight_images_kernel<<<dim3(16, 2 * p.L), 256, 0, st>>>(p);
  *rc = check_cuda(cudaGetLastError(), "umma_weight_images_kernel");
  return ws;
}

int launch_umma_forward(TileParams& p, int passes, float* ws, cudaStream_t st) {
  int rc;
  const int64_t B = p.B;
  int64_t sub = kForwardSub;
  if (p.user_div > 0) sub = std::max<int64_t>(1, sub / p.user_div) * p.user_div;
  carve(p, ws, std::min(B, sub), false, st, &rc);
  if (rc != NCF_OK) return rc;
  const int64_t* user = p.user;
  const int64_t* item = p.item;
  float* logits = p.logits;
  for (int64_t off = 0; off < B; off += sub) {
    p.B = std::min(sub, B - off);
    p.user = (p.user_div > 0) ? user + off / p.user_div : user + off;
    p.item = item + off;
    p.logits = logits + off;
    rc = tower_forward(p, passes, false, st);
    if (rc != NCF_OK) return rc;
  }
  return NCF_OK;
}

int launch_umma_train(TileParams& p, int passes, float* ws, cudaStream_t st) {
  int rc;
  StepTimer timer(st);
  carve(p, ws, p.B, true, st, &rc);
  if (rc != NCF_OK) return rc;
  timer.mark("images");
  rc = tower_forward(p, passes, true, st);
  if (rc != NCF_OK) return rc;
  for (int k = p.L - 1; k >= 0; --k) {
    GemmArgs g{};
    g.K = p.W[k + 1];
    g.n_total = p.W[k];
    g.N = std::min(256, g.n_total);
    g.passes = passes;
    g.layer = k;
    g.a = p.delta[k + 1];
    g.b_img = p.wsplit_b[k];
    if (k > 0) {
      g.out = p.delta[k];
      g.mask_src = p.act[k];
      rc = launch_gemm<EPI_MASK_STORE>(p, g, g.n_total / g.N, st);
    } else {
      rc = launch_gemm<EPI_SCATTER>(p, g, g.n_total / g.N, st);
    }
    if (rc != NCF_OK) return rc;
    timer.mark(k == 2 ? "dgrad2" : k == 1 ? "dgrad1" : k == 0 ? "dgrad0" : "dgrad");
  }
  rc = launch_wgrad(p, passes, st);
  timer.mark("wgrad");
  return rc;
}

}  // namespace ncf
