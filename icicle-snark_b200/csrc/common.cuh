// Shared host-side plumbing for libicicle_b200: error translation, the allocation tracker the
// Rust wrappers assert against, and launch helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>

#include "../../include/icicle_b200.h"

#define ICICLE_UNKNOWN_FALLBACK 12 /* Rust maps index 12 to UnknownError (errors.rs:6-20) */

namespace b200 {

  // cudaError -> eIcicleError, never throwing (the reference throws across extern "C":
  // /root/reference/icicle/backend/cuda/include/gpu-utils/error_handler.h:112-130).
  inline eIcicleError translate(cudaError_t e, eIcicleError fallback)
  {
    if (e == cudaSuccess) return ICICLE_SUCCESS;
    if (e == cudaErrorMemoryAllocation) return ICICLE_OUT_OF_MEMORY;
    if (e == cudaErrorInvalidDevice || e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
      return ICICLE_INVALID_DEVICE;
    if (e == cudaErrorInvalidValue) return ICICLE_INVALID_ARGUMENT;
    return fallback;
  }

#define B200_CUDA(call, fallback)                                                                                      \
  do {                                                                                                                 \
    cudaError_t e__ = (call);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      fprintf(stderr, "[icicle_b200] %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e__), __FILE__, __LINE__);   \
      return b200::translate(e__, fallback);                                                                           \
    }                                                                                                                  \
  } while (0)

#define B200_TRY(expr)                                                                                                 \
  do {                                                                                                                 \
    eIcicleError e__ = (expr);                                                                                         \
    if (e__ != ICICLE_SUCCESS) return e__;                                                                             \
  } while (0)

  // address-range -> device id map (/root/reference/icicle/include/icicle/memory_tracker.h:11-54):
  // anything not in the map is host memory; interior pointers (&d_vec[N..2N]) must resolve.
  class MemoryTracker
  {
  public:
    void add(const void* p, size_t size, int dev)
    {
      std::lock_guard<std::mutex> g(mu_);
      ranges_[(uintptr_t)p] = {size, dev};
    }
    bool remove(const void* p, int* dev)
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = ranges_.find((uintptr_t)p);
      if (it == ranges_.end()) return false;
      if (dev) *dev = it->second.dev;
      ranges_.erase(it);
      return true;
    }
    // returns device id or -1 for host
    int identify(const void* p)
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = ranges_.upper_bound((uintptr_t)p);
      if (it == ranges_.begin()) return -1;
      --it;
      if ((uintptr_t)p < it->first + it->second.size) return it->second.dev;
      return -1;
    }

  private:
    struct R {
      size_t size;
      int dev;
    };
    std::mutex mu_;
    std::map<uintptr_t, R> ranges_;
  };

  MemoryTracker& tracker();
  int active_device();                 // thread-local CUDA ordinal (set by icicle_set_device)
  eIcicleError ensure_device();        // cudaSetDevice(active_device()) + one-time pool setup
  inline cudaStream_t as_stream(icicleStreamHandle s) { return reinterpret_cast<cudaStream_t>(s); }
  int sm_count();

  // stream-ordered scratch from the device's default pool (release threshold = unlimited, so
  // repeated MSMs reuse the same blocks without cudaMalloc churn; the reference pays ~25
  // cudaMallocAsync/FreeAsync pairs per MSM, cuda_msm.cuh:423-482).
  template <class T>
  inline cudaError_t scratch_alloc(T** p, size_t count, cudaStream_t st)
  {
    return cudaMallocAsync((void**)p, count * sizeof(T) + 16, st);
  }
  inline void scratch_free(void* p, cudaStream_t st)
  {
    if (p) cudaFreeAsync(p, st);
  }

  // grid sizing: enough CTAs to cover `work` items, capped to a multiple of the SM count
  inline int grid_for(size_t work, int block, int max_ctas_per_sm = 16)
  {
    size_t g = (work + block - 1) / block;
    size_t cap = (size_t)sm_count() * max_ctas_per_sm;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (int)g;
  }

} // namespace b200
