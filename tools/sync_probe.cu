// Cost of the producer <-> MMA-thread handshake used by tile_umma.cu, in cycles per stage hand-off.
// One CTA: P producer warps, 1 consumer thread, S stages, N iterations, no data movement.
//   variant bit 0: consumer frees the stage with tcgen05.commit (else plain mbarrier.arrive)
//   variant bit 1: producers execute fence.proxy.async before arriving
//   variant bit 2: every producer lane arrives (else one lane per warp)
//   variant bit 3: consumer issues one small tcgen05.mma per stage before the commit
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o sync_probe sync_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_rlx(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rlx(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}

__global__ void probe(int P, int S, int N, int variant, long long* out) {
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ uint32_t tmem_base;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool all_lanes = variant & 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(all_lanes ? P * 32 : P));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == P) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const long long t0 = clock64();
  if (warp < P) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < N; ++i, ++s) {
      if (s == S) { s = 0; ph ^= 1; }
      if (variant & 16) { if (all_lanes) mbar_wait_rlx(&empty[s], ph ^ 1); else { if (lane == 0) mbar_wait_rlx(&empty[s], ph ^ 1); __syncwarp(); } }
      else if (all_lanes) mbar_wait(&empty[s], ph ^ 1);
      else { if (lane == 0) mbar_wait(&empty[s], ph ^ 1); __syncwarp(); }
      if (variant & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (variant & 16) { if (all_lanes) mbar_arrive_rlx(&full[s]); else { __syncwarp(); if (lane == 0) mbar_arrive_rlx(&full[s]); } }
      else if (all_lanes) mbar_arrive(&full[s]);
      else { __syncwarp(); if (lane == 0) mbar_arrive(&full[s]); }
    }
  } else if (lane == 0) {
    // zero operand images: one 128x64x8 tf32 MMA per stage when requested
    uint64_t desc = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16) |
                    (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < N; ++i, ++s) {
      if (s == S) { s = 0; ph ^= 1; }
      if (variant & 16) mbar_wait_rlx(&full[s], ph); else mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;");
      if (variant & 8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(desc), "l"(desc), "r"(idesc), "r"(1));
      if (variant & 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      else if (variant & 16) mbar_arrive_rlx(&empty[s]);
      else mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *out = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == P) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int N = 4096;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int P : {1, 8})
    for (int S : {1, 4})
      for (int variant : {0, 1, 16, 17, 21, 27}) {
        probe<<<1, (P + 1) * 32, 65536>>>(P, S, N, variant, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("P=%d S=%d test_wait=%d commit=%d proxyfence=%d all_lanes=%d mma=%d : %6.0f cycles/hand-off (%s)\n", P, S, (variant >> 4) & 1, variant & 1,
               (variant >> 1) & 1, (variant >> 2) & 1, (variant >> 3) & 1, (double)c / N, cudaGetErrorString(e));
      }
  return 0;
}
