// Probe for the round-2 tcgen05 rewrite of the tile kernel: one CTA, one tcgen05.mma kind::tf32
// chain with accumulators in TMEM, operands in shared memory in the NO-SWIZZLE ("interleaved")
// canonical layout.  It verifies on hardware the three facts the design in DESIGN.md rests on:
//   (1) the descriptor encodings (smem descriptor: start/LBO/SBO/version, instruction descriptor:
//       formats, majors, M/N) as read from the CUTLASS headers;
//   (2) that ONE shared-memory image  [row/8][col/4][row%8][col%4]  (16-byte core-matrix rows) serves
//       both as a K-major operand (rows = M or N, cols = K: forward / dgrad) and as an MN-major
//       operand (rows = K, cols = M or N: wgrad) — so activations need no transposed copy;
//   (3) the accumulator layout in TMEM for M = 128 (row m -> lane m, column n -> column n) as seen
//       by tcgen05.ld.32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run: ./umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;  // D[M][N] = sum_k A[m][k] * B[n][k]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle (layout_type 0), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 0) {
  uint64_t d = (uint64_t)(layout_type & 7) << 61;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 2u << 7;                       // a_format = TF32
  d |= 2u << 10;                      // b_format = TF32
  d |= (uint32_t)a_mn_major << 15;    // 0 = K-major, 1 = MN-major
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// image[(r/8)*(cols/4)*32 + (c/4)*32 + (r%8)*4 + (c%4)] = src[r][c]
__device__ void fill_interleaved(float* img, const float* src, int rows, int cols, int ld) {
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
    const int r = idx / cols, c = idx % cols;
    img[(r >> 3) * (cols >> 2) * 32 + (c >> 2) * 32 + (r & 7) * 4 + (c & 3)] = src[r * ld + c];
  }
}

// Swizzled image of a [rows][cols] fp32 matrix whose rows are 128 bytes (32 floats) wide per atom:
// atoms of 8 rows x 128 B (SWIZZLE_128B: 16-byte chunk index ^= row % 8) or 4 rows x 128 B
// (128B_BASE32B: 32-byte chunk index ^= row % 4).  Atoms are laid out [col/32][row/R] (all row groups
// of one 32-column panel first), so: stride between row groups = R*128 B, between column panels =
// (rows/R) * R*128 B = rows * 128 B.
__device__ void fill_swizzled(float* img, const float* src, int rows, int cols, int ld, int base32) {
  const int R = base32 ? 4 : 8;
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
    const int r = idx / cols, c = idx % cols;
    const int panel = c >> 5, cc = c & 31, rg = r / R, rr = r % R;
    int byte = rr * 128 + cc * 4;
    if (base32) byte = rr * 128 + ((((cc * 4) >> 5) ^ rr) << 5) + ((cc * 4) & 31);
    else byte = rr * 128 + ((((cc * 4) >> 4) ^ rr) << 4) + ((cc * 4) & 15);
    img[(panel * rows * 128 + rg * R * 128 + byte) >> 2] = src[r * ld + c];
  }
}

// mode 0: A [M][K] and B [N][K] as K-major images.
// mode 1: operands given TRANSPOSED in memory, At [K][M] and Bt [K][N], stored with the SAME image
//         rule (rows = K) and consumed as MN-major operands.
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int mode,
                                                    int* status) {
  extern __shared__ __align__(1024) float smem[];
  float* a_img = smem;                 // M*K floats
  float* b_img = smem + M * K;         // N*K floats
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const bool a_mn = (mode == 1 || mode == 2 || mode == 3 || mode == 5);
  const bool b_mn = (mode == 1 || mode == 2 || mode == 4 || mode == 6);
  const bool swp = (mode == 2 || mode == 5 || mode == 6);
  if (mode == 7 || mode == 10 || mode == 11) {            // K-major, SWIZZLE_128B: rows = M/N, cols = K (one 32-column panel)
    fill_swizzled(a_img, A, M, K, K, 0);
    fill_swizzled(b_img, B, N, K, K, 0);
  } else if (mode == 8 || mode == 9) {     // MN-major: rows = K, cols = M/N; 8: SWIZZLE_128B, 9: 128B_BASE32B
    fill_swizzled(a_img, A, K, M, M, mode == 9);
    fill_swizzled(b_img, B, K, N, N, mode == 9);
  } else {
    if (!a_mn) fill_interleaved(a_img, A, M, K, K); else fill_interleaved(a_img, A, K, M, M);
    if (!b_mn) fill_interleaved(b_img, B, N, K, K); else fill_interleaved(b_img, B, K, N, N);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of the operand images must be visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (mode == 11) {
    // A operand in tensor memory: thread = row m (lane 32*warp + lane), element (m, k) at column 64 + k
    const int row = warp * 32 + lane;
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(A[row * K + k]);
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(M, N, a_mn || mode == 8 || mode == 9, b_mn || mode == 8 || mode == 9);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      if (mode == 7 || mode == 10 || mode == 11) {
        // K-major SW128: 8-row groups 1024 B apart (SBO); the k-step advances 32 B inside the 128-B row
        da = make_desc(smem_u32(a_img) + ks * 32, 16, 1024, 2);
        db = make_desc(smem_u32(b_img) + ks * 32, 16, 1024, 2);
      } else if (mode == 8) {
        // MN-major SW128: image rows = K (8 per atom); LBO = stride between 32-element MN panels,
        // SBO = stride between 8-row K groups
        da = make_desc(smem_u32(a_img) + ks * 1024, K * 128, 1024, 2);
        db = make_desc(smem_u32(b_img) + ks * 1024, K * 128, 1024, 2);
      } else if (mode == 9) {
        // MN-major 128B_BASE32B: 4-row K atoms (512 B); one MMA (K=8) spans two of them
        da = make_desc(smem_u32(a_img) + ks * 1024, K * 128, 512, 1);
        db = make_desc(smem_u32(b_img) + ks * 1024, K * 128, 512, 1);
      } else if (!a_mn) da = make_desc(smem_u32(a_img) + ks * 2 * 128, 128, (K / 4) * 128);
      else if (!swp) da = make_desc(smem_u32(a_img) + ks * (M / 4) * 128, (M / 4) * 128, 128);
      else da = make_desc(smem_u32(a_img) + ks * (M / 4) * 128, 128, (M / 4) * 128);
      if (mode >= 7) {
      } else if (!b_mn) db = make_desc(smem_u32(b_img) + ks * 2 * 128, 128, (K / 4) * 128);
      else if (!swp) db = make_desc(smem_u32(b_img) + ks * (N / 4) * 128, (N / 4) * 128, 128);
      else db = make_desc(smem_u32(b_img) + ks * (N / 4) * 128, 128, (N / 4) * 128);
      const uint32_t acc = ks > 0;
      if (mode == 11) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(tmem), "r"(tmem + 64 + ks * 8), "l"(db), "r"(idesc), "r"(acc));
        continue;
      }
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // bounded wait on the MMA-completion barrier (phase 0)
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
  }
  if (!done) { if (threadIdx.x == 0) *status = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (done) {
    // warp w may read TMEM lanes [32w, 32w+32): thread = row, 16 columns per load
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

static float tf32_exact(float x) {  // keep 10 mantissa bits so the reference is exact
  uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x;
}

int main() {
  std::vector<float> A(M * K), B(N * K), At(K * M), Bt(K * N), ref(M * N), out(M * N);
  srand(1);
  for (auto& v : A) v = tf32_exact((rand() % 2001 - 1000) / 512.f);
  for (auto& v : B) v = tf32_exact((rand() % 2001 - 1000) / 512.f);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) At[k * M + m] = A[m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bt[k * N + n] = B[n * K + k];
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
    ref[m * N + n] = (float)s;
  }
  float *dA, *dB, *dD; int* dS;
  cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K);
  cudaMalloc(&dD, sizeof(float) * M * N); cudaMalloc(&dS, sizeof(int));
  const size_t smem = sizeof(float) * (M * K + N * K);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rc = 0;
  for (int mode = 0; mode < 12; ++mode) {
    // mode 10: operands carry 0.75 ulp(tf32) of extra low mantissa bits; the result equals the exact
    // one iff the tensor core TRUNCATES fp32 -> tf32 (round-to-nearest would move every operand up)
    std::vector<float> Alow(A), Blow(B);
    for (auto& v : Alow) { uint32_t u; memcpy(&u, &v, 4); u |= 0x1800u; memcpy(&v, &u, 4); }
    for (auto& v : Blow) { uint32_t u; memcpy(&u, &v, 4); u |= 0x1800u; memcpy(&v, &u, 4); }
    cudaMemcpy(dA, mode == 10 ? Alow.data() : (mode==1||mode==2||mode==3||mode==5||mode==8||mode==9) ? At.data() : A.data(), sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, mode == 10 ? Blow.data() : (mode==1||mode==2||mode==4||mode==6||mode==8||mode==9) ? Bt.data() : B.data(), sizeof(float) * N * K, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof(float) * M * N); cudaMemset(dS, 0, sizeof(int));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > worst) worst = d; if (d > 1e-3) ++bad; }
    printf("mode %d (%s operands): cuda=%s barrier_timeout=%d max_abs_err=%.3g mismatches=%d/%d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
           mode, mode ? "variant" : "K-major", cudaGetErrorString(e), st, worst, bad, M * N,
           out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    if (mode == 0 && (e != cudaSuccess || st || bad)) rc = 1;
    if (e != cudaSuccess) break;
  }
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  return rc;
}
