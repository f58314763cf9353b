// Peer-memory plumbing for single-node data parallelism: buffers that other ranks' kernels read and
// write over NVLink (ncf_adam_p2p in optim.cu).  The reference is single-process; this is new design
// (SURVEY.md section 8e).
//
// CUDA IPC hands out whole allocations, so a shared buffer has to be its own cudaMalloc allocation
// rather than a slice of the framework's caching allocator: ncf_peer_alloc / ncf_peer_free are the one
// place where the library owns device memory, and only on the caller's explicit request.
#include "common.cuh"

extern "C" int ncf_peer_alloc(int64_t bytes, void** dev_ptr_out) {
  NCF_REQUIRE(bytes > 0 && dev_ptr_out, "ncf_peer_alloc: bad argument");
  void* p = nullptr;
  NCF_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    cudaFree(p);
    NCF_CUDA(e);
  }
  *dev_ptr_out = p;
  return NCF_OK;
}

extern "C" int ncf_peer_free(void* dev_ptr) {
  if (dev_ptr) NCF_CUDA(cudaFree(dev_ptr));
  return NCF_OK;
}

// handle64: 64 bytes (cudaIpcMemHandle_t), to be sent to the other ranks by any host channel
extern "C" int ncf_ipc_export(const void* dev_ptr, void* handle64) {
  NCF_REQUIRE(dev_ptr && handle64, "ncf_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  NCF_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle64, &h, sizeof(h));
  return NCF_OK;
}

// Maps another process's allocation into this process, accessible from the current device.
extern "C" int ncf_ipc_open(const void* handle64, void** dev_ptr_out) {
  NCF_REQUIRE(handle64 && dev_ptr_out, "ncf_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  NCF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return NCF_OK;
}

extern "C" int ncf_ipc_close(void* dev_ptr) {
  if (dev_ptr) NCF_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return NCF_OK;
}

// ---- rank barrier over peer memory ---------------------------------------------------------------------------------
// One CTA per rank: the rank publishes its arrival (the number of this barrier, counted on the device so that
// the kernel can live in a CUDA graph) into slot `rank` of EVERY rank's flag array with a system-scope
// release store over NVLink, then waits until every slot of its OWN array has reached that number.  The
// release / acquire pair orders everything this GPU wrote before the barrier (earlier kernels on the stream)
// against what the peers read after it.  ~2 NVLink latencies instead of a NCCL launch; the wait is bounded:
// a rank that never arrives ends in a trap (launch failure), not in a hung GPU.
namespace {

struct BarrierPeers {
  uint32_t* flags[8];
};

__global__ void peer_barrier_kernel(const __grid_constant__ BarrierPeers peers, int world, int rank,
                                    uint32_t* __restrict__ epoch_counter) {
  __shared__ uint32_t epoch_sm;
  if (threadIdx.x == 0) {
    const uint32_t e = *epoch_counter + 1;
    *epoch_counter = e;
    epoch_sm = e;
  }
  __syncthreads();
  const uint32_t epoch = epoch_sm;
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.flags[t] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = peers.flags[rank] + t;
    bool ok = false;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spin = 0;; ++spin) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) { ok = true; break; }
      if ((spin & 1023u) == 1023u) {     // bounded wait: 30 s of wall clock, then a trap instead of a hung GPU
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 30ull * 1000000000ull) break;
      }
    }
    if (!ok) __trap();
    __threadfence_system();
  }
  __syncthreads();
}

}  // namespace

extern "C" int ncf_peer_barrier(void* const* flag_peers, int32_t world, int32_t rank, uint32_t* epoch_counter,
                                void* stream) {
  NCF_REQUIRE(flag_peers && epoch_counter, "ncf_peer_barrier: null pointer");
  NCF_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "ncf_peer_barrier: world must be 1..8 (one node)");
  BarrierPeers p{};
  for (int r = 0; r < world; ++r) {
    NCF_REQUIRE(flag_peers[r] != nullptr && ((uintptr_t)flag_peers[r] & 3) == 0, "ncf_peer_barrier: flag array %d is NULL or misaligned", r);
    p.flags[r] = reinterpret_cast<uint32_t*>(flag_peers[r]);
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, world, rank, epoch_counter);
  NCF_LAUNCH_CHECK("peer_barrier_kernel");
  return NCF_OK;
}
