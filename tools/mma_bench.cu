// Microbenchmark: issue rate of legacy warp-level mma.sync on sm_100a (tf32 m16n8k8, bf16 m16n8k16)
// and of FFMA, to decide which pipe the fp32-parity tower should run on.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int ILP>
__global__ void tf32_kernel(float* out, int iters) {
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void bf16_kernel(float* out, int iters) {
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// realistic operand pattern: 2 A fragments x 4 B fragments, all distinct registers, refreshed by ALU ops
__global__ void tf32_real_kernel(float* out, int iters) {
  float c[2][4][4];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[i][j][q] = 0.f;
  unsigned a[2][4], b[4][2];
  for (int i = 0; i < 2; ++i) for (int q = 0; q < 4; ++q) a[i][q] = threadIdx.x * 7 + i * 4 + q;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 2; ++q) b[j][q] = threadIdx.x * 3 + j * 2 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][j][0]), "+f"(c[i][j][1]), "+f"(c[i][j][2]), "+f"(c[i][j][3])
                     : "r"(a[i][0]), "r"(a[i][1]), "r"(a[i][2]), "r"(a[i][3]), "r"(b[j][0]), "r"(b[j][1]));
#pragma unroll
    for (int i = 0; i < 2; ++i) for (int q = 0; q < 4; ++q) a[i][q] += 0x2000;
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 2; ++q) b[j][q] ^= 0x4000;
  }
  float s = 0; for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[i][j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
  float c[16];
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3f + i;
  float a = 1.0001f, b = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32, blocks = sms;
    float ms = time_ms([&] { tf32_kernel<8><<<blocks, threads>>>(out, iters); });
    double flop = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * blocks;
    printf("tf32 m16n8k8  warps/SM=%2d ILP=8: %.1f TFLOP/s (%.3f ms)\n", warps, flop / ms / 1e9, ms);
    ms = time_ms([&] { bf16_kernel<8><<<blocks, threads>>>(out, iters); });
    flop = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * blocks;
    printf("bf16 m16n8k16 warps/SM=%2d ILP=8: %.1f TFLOP/s (%.3f ms)\n", warps, flop / ms / 1e9, ms);
  }
  for (int warps : {8, 16}) {
    float msr = time_ms([&] { tf32_real_kernel<<<sms, warps * 32>>>(out, iters); });
    printf("tf32 realistic operands warps/SM=%2d: %.1f TFLOP/s\n", warps, 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * sms / msr / 1e9);
  }
  float ms = time_ms([&] { ffma_kernel<<<sms * 8, 256>>>(out, iters); });
  printf("ffma: %.1f TFLOP/s\n", 2.0 * 16 * iters * 256.0 * sms * 8 / ms / 1e9);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
