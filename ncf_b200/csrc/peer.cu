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
