// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/yocto_hair_b200.h"

namespace yh {

std::string& last_error_storage();
yh_status    fail(yh_status code, const char* fmt, ...);
// true when a CUDA device is usable; otherwise sets YH_ERR_NO_DEVICE
bool         require_device();

#define YH_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t err_ = (call);                                                              \
    if (err_ != cudaSuccess)                                                                \
      return yh::fail(YH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err_), \
          __FILE__, __LINE__);                                                              \
  } while (0)

#define YH_REQUIRE_DEVICE()                          \
  do {                                               \
    if (!yh::require_device()) return YH_ERR_NO_DEVICE; \
  } while (0)

// raises the release threshold of the current device's default memory pool (once per device)
void pool_keep_memory();
// YH_NO_POOL=1: cudaMalloc / cudaFree instead of the stream-ordered pool (diagnostics: compute-sanitizer then sees
// the bounds of every buffer)
bool plain_allocations();

// device buffer with RAII (host-side owner)
template <typename T>
struct dev_buffer {
  T*     ptr   = nullptr;
  size_t count = 0;
  dev_buffer() = default;
  dev_buffer(const dev_buffer&)            = delete;
  dev_buffer& operator=(const dev_buffer&) = delete;
  ~dev_buffer() { release(); }
  // Allocations come from the device's default memory pool (stream-ordered allocator on the
  // legacy stream, followed by a synchronise so that any stream may use the memory); the pool
  // keeps freed blocks (see pool_keep_memory()), so re-creating a scene or state does not pay
  // cudaMalloc / cudaFree's map / unmap cost again (they were seen to take 100s of ms).
  void release() {
    if (ptr) {
      if (plain_allocations()) cudaFree(ptr);
      else cudaFreeAsync(ptr, 0);
    }
    ptr   = nullptr;
    count = 0;
  }
  cudaError_t alloc(size_t n) {
    release();
    count = n;
    if (n == 0) return cudaSuccess;
    if (plain_allocations()) return cudaMalloc((void**)&ptr, n * sizeof(T));  // exact bounds for compute-sanitizer
    pool_keep_memory();
    cudaError_t e = cudaMallocAsync((void**)&ptr, n * sizeof(T), 0);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(0);
  }
  cudaError_t upload(const T* host, size_t n) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess || n == 0) return e;
    e = cudaMemcpy(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    // cudaMemcpy from PAGEABLE host memory returns once the data sits in the driver's staging buffers; the DMA into
    // `ptr` is ordered only against the legacy stream.  The consumers here launch on cudaStreamNonBlocking streams,
    // which do not wait for it (seen: k_init_slots reading the pixel map of a 720 x 405 state before it had landed),
    // so drain the legacy stream before anyone may use the buffer.
    return cudaStreamSynchronize(0);
  }
  size_t bytes() const { return count * sizeof(T); }
};

// number of SMs of the current device (148 on B200), cached
int sm_count();

}  // namespace yh
