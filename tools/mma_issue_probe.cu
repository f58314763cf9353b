// How fast can one thread feed tcgen05.mma kind::tf32 (K = 8 per instruction)?  Cycles per MMA for a
// chain of R instructions, by N, operand source of A (shared memory / tensor memory) and the number of
// distinct accumulators the chain rotates over (1 = every MMA depends on the previous one).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_issue_probe mma_issue_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(int N, int ts, int nacc, int R, long long* out, int W, int same_acc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(W));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if ((threadIdx.x & 31) == 0 && warp < W) {
    const uint32_t tmem = tmem_base + (same_acc ? 0 : warp * N);
    const uint64_t da = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16) |
                        (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
    const uint64_t db = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16) |
                        (uint64_t)((smem_u32(smem + 16384) >> 4) & 0x3fff);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    int acc = 0;
#pragma unroll 4
    for (int i = 0; i < R; ++i) {
      const uint32_t d = tmem + acc * N;  // accumulators side by side; A operand (TS) in the last 32 columns
      if (++acc == nacc) acc = 0;
      if (ts)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tmem_base + 480 + (i & 3) * 8), "l"(db), "r"(idesc), "r"(1));
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(idesc), "r"(1));
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int R = 96;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128})
      for (int W : {1, 2, 3, 4})
        for (int same : {0, 1}) {
          if (!same && W * N > 448) continue;
          probe<<<1, 128, 64 * 1024>>>(N, ts, 1, R, d, W, same);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("A from %s  N=%3d  issuing warps=%d %s accumulator: %5.1f cyc per MMA overall (floor N/2 = %d) %s\n",
                 ts ? "TMEM" : "smem", N, W, same ? "one shared" : "own", (double)h[1] / (R * W), N / 2, cudaGetErrorString(e));
        }
  return 0;
}
