// Fixed-base batch scalar multiplication: out[i] = k_i * G for the BN254 G1 / G2 generators.
// The synthetic trusted-setup generator (tools/synth.py) uses it to build valid .zkey
// files of benchmark size without snarkjs, and the tests use it for known-discrete-log MSM checks at
// sizes the CPU oracle cannot reach.  8-bit fixed windows: 32 tables of 255 affine
// multiples (built once per process on the device), 32 mixed adds + one inversion per output point.
#include <mutex>

#include "host_math.h"
#include "msm.cuh"

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  template <class T>
  __device__ __forceinline__ void st16(T* p, const T& v)
  {
    constexpr int NQ = sizeof(T) / 16;
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* w = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      q[i] = w[i];
  }
  template <class T>
  __device__ __forceinline__ T ld16(const T* p)
  {
    constexpr int NQ = sizeof(T) / 16;
    T r;
    uint4* w = reinterpret_cast<uint4*>(&r);
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      w[i] = q[i];
    return r;
  }

  // thread w builds table[w][d-1] = d * 2^(8w) * G, d = 1..255 (affine, Montgomery)
  template <class F>
  __global__ void fixed_base_table_kernel(Affine<F> gen, Affine<F>* table)
  {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    XYZZ<F> p = XYZZ<F>::from_affine(gen);
    for (int k = 0; k < 8 * w; ++k)
      p = p.dbl();
    Affine<F> base = p.to_affine();
    XYZZ<F> acc = XYZZ<F>::from_affine(base);
    for (int d = 1; d <= 255; ++d) {
      Affine<F> a = acc.to_affine();
      st16(table + w * 255 + (d - 1), a);
      acc = XYZZ<F>::from_affine(a);
      acc.madd(base);
    }
  }

  template <class F>
  __global__ void __launch_bounds__(128)
    fixed_base_mul_kernel(const Fr* k, size_t n, const Affine<F>* table, bool out_mont, Affine<F>* out)
  {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      Fr s = ld16(k + i);
      XYZZ<F> acc = XYZZ<F>::inf();
      for (int w = 0; w < 32; ++w) {
        uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 0xff;
        if (d) acc.madd(ld16(table + w * 255 + (d - 1)));
      }
      Affine<F> a = acc.to_affine();
      if (!out_mont) a = {F::from_mont(a.x), F::from_mont(a.y)};
      st16(out + i, a);
    }
  }

  static std::mutex g_tab_mu;
  static void* g_tables[64][2] = {};

  template <class F>
  static eIcicleError get_table(bool g2, const Affine<F>& gen, Affine<F>** out)
  {
    int dev = active_device();
    if (dev < 0 || dev >= 64) return ICICLE_INVALID_DEVICE;
    std::lock_guard<std::mutex> g(g_tab_mu);
    if (!g_tables[dev][g2]) {
      Affine<F>* t = nullptr;
      B200_CUDA(cudaMalloc((void**)&t, 32 * 255 * sizeof(Affine<F>)), ICICLE_ALLOCATION_FAILED);
      B200_LAUNCH(fixed_base_table_kernel<F>, 1, 32, 0, 0, gen, t);
      B200_CUDA(cudaDeviceSynchronize(), ICICLE_SYNCHRONIZATION_FAILED);
      g_tables[dev][g2] = t;
    }
    *out = (Affine<F>*)g_tables[dev][g2];
    return ICICLE_SUCCESS;
  }

  template <class F>
  static eIcicleError fixed_base_mul(const void* k, size_t n, bool g2, const Affine<F>& gen, bool out_mont, void* out)
  {
    if (!k || !out) return ICICLE_INVALID_POINTER;
    B200_TRY(ensure_device());
    if (n == 0) return ICICLE_SUCCESS;
    Affine<F>* table = nullptr;
    B200_TRY(get_table<F>(g2, gen, &table));
    StagedIn K;
    StagedOut O;
    B200_TRY(K.init(k, n * 32, false, 0));
    B200_TRY(O.init(out, n * sizeof(Affine<F>), false, 0));
    B200_LAUNCH(fixed_base_mul_kernel<F>, grid_for(n, 128, 8), 128, 0, 0, (const Fr*)K.dev, n, table, out_mont, (Affine<F>*)O.dev);
    B200_CUDA(cudaGetLastError(), ICICLE_UNKNOWN_FALLBACK);
    B200_TRY(O.finish(0));
    K.release(0);
    B200_CUDA(cudaStreamSynchronize(0), ICICLE_SYNCHRONIZATION_FAILED);
    return ICICLE_SUCCESS;
  }

} // namespace b200

using namespace b200;

// out[i] = k_i * G1 (g2 == 0, 64 B affine) or k_i * G2 (g2 != 0, 128 B affine); k standard form, host or
// icicle_malloc'd device memory; output in Montgomery form (as .zkey stores points) or standard form.
extern "C" eIcicleError b200_fixed_base_mul(const bn254_scalar_t* k, uint64_t n, int g2, int out_montgomery, void* out)
{
  if (g2) return fixed_base_mul<Fq2>(k, n, true, g2_generator_mont(), out_montgomery != 0, out);
  return fixed_base_mul<Fq>(k, n, false, g1_generator_mont(), out_montgomery != 0, out);
}
