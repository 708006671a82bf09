// Host<->device operand staging for the op-level C ABI (the reference stages host operands with
// cudaMallocAsync+cudaMemcpyAsync too: cuda_vec_ops.cu:17-54) and 128-bit load/store helpers.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace b200 {

  // An operand is device-resident if the caller's flag says so OR the pointer lies inside an
  // icicle_malloc'd range (the tracker is authoritative: SURVEY App. C, field.rs:379-398 quirk).
  inline bool is_device_ptr(const void* p, bool flag) { return flag || tracker().identify(p) >= 0; }

  struct StagedIn {
    const void* dev = nullptr;
    void* tmp = nullptr;
    eIcicleError init(const void* p, size_t bytes, bool flag, cudaStream_t st)
    {
      if (is_device_ptr(p, flag)) {
        dev = p;
        return ICICLE_SUCCESS;
      }
      B200_CUDA(cudaMallocAsync(&tmp, bytes ? bytes : 16, st), ICICLE_ALLOCATION_FAILED);
      B200_CUDA(cudaMemcpyAsync(tmp, p, bytes, cudaMemcpyHostToDevice, st), ICICLE_COPY_FAILED);
      dev = tmp;
      return ICICLE_SUCCESS;
    }
    void release(cudaStream_t st)
    {
      if (tmp) cudaFreeAsync(tmp, st);
      tmp = nullptr;
    }
  };

  struct StagedOut {
    void* dev = nullptr;
    void* tmp = nullptr;
    void* host = nullptr;
    size_t bytes = 0;
    eIcicleError init(void* p, size_t nbytes, bool flag, cudaStream_t st)
    {
      bytes = nbytes;
      if (is_device_ptr(p, flag)) {
        dev = p;
        return ICICLE_SUCCESS;
      }
      host = p;
      B200_CUDA(cudaMallocAsync(&tmp, bytes ? bytes : 16, st), ICICLE_ALLOCATION_FAILED);
      dev = tmp;
      return ICICLE_SUCCESS;
    }
    // enqueue the copy back; valid on the host after the stream is synchronised
    eIcicleError finish(cudaStream_t st)
    {
      if (tmp) {
        B200_CUDA(cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, st), ICICLE_COPY_FAILED);
        cudaFreeAsync(tmp, st);
        tmp = nullptr;
      }
      return ICICLE_SUCCESS;
    }
  };

#if defined(__CUDACC__)
  // 32 B field element as two 128-bit transactions
  template <class F>
  __device__ __forceinline__ F ld_fp(const F* p)
  {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    F r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  // same, but through the coherent path (for buffers written earlier in the same kernel)
  template <class F>
  __device__ __forceinline__ F ld_fp_coherent(const F* p)
  {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    F r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  template <class F>
  __device__ __forceinline__ void st_fp(F* p, const F& x)
  {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
  __device__ __forceinline__ Fr ld_fr(const Fr* p) { return ld_fp<Fr>(p); }
  __device__ __forceinline__ void st_fr(Fr* p, const Fr& x) { st_fp<Fr>(p, x); }
#endif

} // namespace b200
