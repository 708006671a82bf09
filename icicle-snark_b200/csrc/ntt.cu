// Fr NTT: domain (twiddle table resident in HBM), Stockham pass kernel, and the bn254_ntt* symbols.
// See ntt.cuh for the design table and the reference locations this replaces.
#include <mutex>
#include <vector>

#include "ntt.cuh"
#include "staging.cuh"
#include "msm.cuh" // g_launches

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  // log2 of the largest per-pass radix and log2 of the elements per CTA tile (32 B each, in shared memory).  Radix-4
  // stages (two butterfly levels per shared-memory round trip and per barrier) are the default; B200_NTT_VARIANT=0 selects
  // the radix-2 stage loop, B200_NTT_KMAX / B200_NTT_TILE_LOG the pass structure (tuning knobs, read once).
  struct NttTuning {
    int kmax, tile_log, variant;
  };
  static const NttTuning& ntt_tuning()
  {
    static const NttTuning t = [] {
      NttTuning v{8, 10, 1};
      if (const char* e = getenv("B200_NTT_KMAX")) v.kmax = atoi(e);
      if (const char* e = getenv("B200_NTT_TILE_LOG")) v.tile_log = atoi(e);
      if (const char* e = getenv("B200_NTT_VARIANT")) v.variant = atoi(e);
      if (v.kmax < 4) v.kmax = 4; // at most 7 passes up to 2^28 (plan_passes)
      if (v.kmax > 12) v.kmax = 12;
      if (v.tile_log < v.kmax) v.tile_log = v.kmax;
      if (v.tile_log > 12) v.tile_log = 12; // 4096 x 32 B = 128 KiB + the twiddle table
      return v;
    }();
    return t;
  }

  // ------------------------------------------------------------------------------------------------ domain
  static NttDomain g_domains[64];
  static std::mutex g_dom_mu;

  const NttDomain* ntt_domain()
  {
    int dev = active_device();
    if (dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> g(g_dom_mu);
    return g_domains[dev].max_log >= 0 ? &g_domains[dev] : nullptr;
  }

  Fr host_omega(int logn)
  {
    // rou of order 2^28 (bn254_scalar.h:68-69), squared down (modular_arithmetic.h:61-73)
    static const uint32_t rou[8] = {0x725b19f0, 0x9bd61b6e, 0x41112ed4, 0x402d111e,
                                    0x8ef62abc, 0x00e0a7eb, 0xa58a7e85, 0x2a3c09f0};
    Fr w;
    memcpy(w.v, rou, 32);
    w = Fr::to_mont(w);
    for (int i = 28; i > logn; --i)
      w = w.sqr();
    return Fr::from_mont(w);
  }

  struct PowTable {
    Fr pw[28]; // root^(2^j), Montgomery
  };

  // tw[i] = root^i for i in [0, 2^max_log]; each thread multiplies the table entries of its index bits
  static __global__ void __launch_bounds__(256) twiddle_fill_kernel(PowTable t, int max_log, Fr* tw)
  {
    size_t n = ((size_t)1 << max_log) + 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      Fr acc = Fr::one();
      for (int b = 0; b < max_log; ++b)
        if ((i >> b) & 1) acc = acc * t.pw[b];
      st_fr(tw + i, acc); // i == 2^max_log has no table bit set: root^(2^max_log) == 1
    }
  }

  eIcicleError ntt_init_domain_host(const Fr& root_std, cudaStream_t st)
  {
    B200_TRY(ensure_device());
    int dev = active_device();
    if (dev < 0 || dev >= 64) return ICICLE_INVALID_DEVICE;
    std::lock_guard<std::mutex> g(g_dom_mu);
    NttDomain& d = g_domains[dev];
    if (d.max_log >= 0) return ICICLE_SUCCESS; // already initialised: no-op like ntt.cuh:452
    // order of the root: square until one (ntt.cuh:459-467); must be a 2^k-th root of unity, k <= 28
    PowTable t;
    Fr w = Fr::to_mont(root_std);
    int max_log = 0;
    while (w != Fr::one()) {
      if (max_log >= 28) return ICICLE_INVALID_ARGUMENT;
      t.pw[max_log++] = w;
      w = w.sqr();
    }
    size_t n = ((size_t)1 << max_log) + 1;
    Fr* tw = nullptr;
    B200_CUDA(cudaMalloc((void**)&tw, n * sizeof(Fr)), ICICLE_ALLOCATION_FAILED);
    B200_LAUNCH(twiddle_fill_kernel, grid_for(n, 256, 8), 256, 0, st, t, max_log, tw);
    B200_CUDA(cudaGetLastError(), ICICLE_UNKNOWN_FALLBACK);
    B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED);
    d.max_log = max_log;
    d.tw = tw;
    d.root_std = root_std;
    return ICICLE_SUCCESS;
  }

  static eIcicleError ntt_release_domain_host()
  {
    int dev = active_device();
    if (dev < 0 || dev >= 64) return ICICLE_INVALID_DEVICE;
    std::lock_guard<std::mutex> g(g_dom_mu);
    NttDomain& d = g_domains[dev];
    if (d.max_log < 0) return ICICLE_SUCCESS;
    cudaDeviceSynchronize();
    cudaFree(d.tw);
    d.tw = nullptr;
    d.max_log = -1;
    return ICICLE_SUCCESS;
  }

  // ------------------------------------------------------------------------------------------------ pass kernel
  struct NttPassArgs {
    const Fr* in;
    Fr* out;
    const Fr* tw;          // w^i, i in [0, 2^log_nmax]
    const Fr* post_table;  // nullable; applied on the last pass
    Fr scale;              // 1/N (Montgomery) when has_scale
    size_t batch_stride;   // elements between consecutive transforms (rows batch) or 1 (columns batch)
    int estride;           // 1 (rows batch) or batch (columns batch)
    int logn, ns_log, k, m, log_nmax;
    int has_scale, last;
  };

  __device__ __forceinline__ Fr lds_fr(const uint4* s, int idx)
  {
    uint4 lo = s[2 * idx], hi = s[2 * idx + 1];
    Fr r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  __device__ __forceinline__ void sts_fr(uint4* s, int idx, const Fr& x)
  {
    s[2 * idx] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    s[2 * idx + 1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }

  // One Stockham pass of radix R = 2^k with Ns = 2^ns_log already-transformed points per group:
  //   v_t   = in[j + t*N/R] * w_N^{(j mod Ns) * t * N/(Ns*R)}          t in [0,R)
  //   V     = DFT_R(v)                                                  (radix-2 DIF in shared memory)
  //   out[(j div Ns)*Ns*R + (j mod Ns) + q*Ns] = V_q
  // A CTA owns 2^m consecutive j (coalesced 2^m*32 B rows on both sides). Hot loop #3.
  template <bool INV>
  __global__ void __launch_bounds__(256) ntt_pass_kernel(NttPassArgs a)
  {
    extern __shared__ uint4 smem[];
    const int R = 1 << a.k, M = 1 << a.m;
    uint4* s = smem;                 // R*M elements
    uint4* itw = smem + 2 * (R * M); // R/2 internal twiddles w_R^i
    const int nthreads = blockDim.x, tid = threadIdx.x;
    const uint32_t nmax = 1u << a.log_nmax;
    const int log_nr = a.logn - a.k; // log2(N/R)
    const uint32_t j0 = blockIdx.x << a.m;
    const uint32_t ns_mask = (1u << a.ns_log) - 1;
    const Fr* in = a.in + (size_t)blockIdx.y * a.batch_stride;
    Fr* out = a.out + (size_t)blockIdx.y * a.batch_stride;

    for (int i = tid; i < R / 2; i += nthreads) {
      uint32_t e = (uint32_t)i << (a.log_nmax - a.k);
      if (INV && e) e = nmax - e;
      sts_fr(itw, i, ld_fr(a.tw + e));
    }
    const int tw_shift = a.log_nmax - a.ns_log - a.k;
    for (int e = tid; e < R * M; e += nthreads) {
      uint32_t t = e >> a.m, jj = e & (M - 1);
      uint32_t j = j0 + jj;
      Fr v = ld_fp_coherent(in + ((size_t)j + ((size_t)t << log_nr)) * a.estride);
      if (a.ns_log > 0) {
        uint32_t x = ((j & ns_mask) * t) << tw_shift;
        if (x) {
          if (INV) x = nmax - x;
          v = v * ld_fr(a.tw + x);
        }
      }
      sts_fr(s, e, v);
    }
    __syncthreads();

    const int nbf = (R / 2) << a.m;
    for (int st = a.k - 1; st >= 0; --st) {
      const uint32_t half = 1u << st;
      for (int b = tid; b < nbf; b += nthreads) {
        uint32_t jj = b & (M - 1), bb = b >> a.m;
        uint32_t lo = bb & (half - 1);
        uint32_t t = ((bb >> st) << (st + 1)) | lo;
        int i0 = (t << a.m) + jj, i1 = ((t + half) << a.m) + jj;
        Fr x = lds_fr(s, i0), y = lds_fr(s, i1);
        Fr d = x - y;
        if (st > 0 && lo) d = d * lds_fr(itw, lo << (a.k - 1 - st)); // last stage: all twiddles are 1
        sts_fr(s, i0, x + y);
        sts_fr(s, i1, d);
      }
      __syncthreads();
    }

    for (int e = tid; e < R * M; e += nthreads) {
      uint32_t q, jj;
      if (a.ns_log == 0) { // first pass: the tile's outputs are one contiguous run, q fastest
        jj = e >> a.k;
        q = e & (R - 1);
      } else {
        q = e >> a.m;
        jj = e & (M - 1);
      }
      uint32_t j = j0 + jj;
      size_t idx = ((size_t)(j >> a.ns_log) << (a.ns_log + a.k)) + (j & ns_mask) + ((size_t)q << a.ns_log);
      uint32_t qr = a.k ? __brev(q) >> (32 - a.k) : 0;
      Fr v = lds_fr(s, (qr << a.m) + jj);
      if (a.last) {
        if (a.post_table)
          v = v * ld_fr(a.post_table + idx);
        else if (a.has_scale)
          v = v * a.scale;
      }
      st_fr(out + idx * a.estride, v);
    }
  }

  template <bool INV>
  __global__ void __launch_bounds__(512) ntt_pass4_kernel(NttPassArgs a)
  {
    extern __shared__ uint4 smem[];
    const int R = 1 << a.k, M = 1 << a.m;
    uint4* s = smem;                 // R*M elements
    uint4* itw = smem + 2 * (R * M); // R/2 internal twiddles w_R^i
    const int nthreads = blockDim.x, tid = threadIdx.x;
    const uint32_t nmax = 1u << a.log_nmax;
    const int log_nr = a.logn - a.k; // log2(N/R)
    const uint32_t j0 = blockIdx.x << a.m;
    const uint32_t ns_mask = (1u << a.ns_log) - 1;
    const Fr* in = a.in + (size_t)blockIdx.y * a.batch_stride;
    Fr* out = a.out + (size_t)blockIdx.y * a.batch_stride;

    for (int i = tid; i < R / 2; i += nthreads) {
      uint32_t e = (uint32_t)i << (a.log_nmax - a.k);
      if (INV && e) e = nmax - e;
      sts_fr(itw, i, ld_fr(a.tw + e));
    }
    const int tw_shift = a.log_nmax - a.ns_log - a.k;
    for (int e = tid; e < R * M; e += nthreads) {
      uint32_t t = e >> a.m, jj = e & (M - 1);
      uint32_t j = j0 + jj;
      Fr v = ld_fp_coherent(in + ((size_t)j + ((size_t)t << log_nr)) * a.estride);
      if (a.ns_log > 0) {
        uint32_t x = ((j & ns_mask) * t) << tw_shift;
        if (x) {
          if (INV) x = nmax - x;
          v = v * ld_fr(a.tw + x);
        }
      }
      sts_fr(s, e, v);
    }
    __syncthreads();

    // radix-2 DIF levels taken two at a time: the four elements {t0, t0+q, t0+h, t0+h+q} (h = 2^st, q = h/2) go through
    // level st (pairs h apart, twiddles w^e0 and w^(e0 + R/4)) and level st-1 (pairs q apart, twiddle w^(2 e0)) in
    // registers - one shared-memory round trip and one barrier per two levels.  An odd k starts with one radix-2 level.
    int st = a.k - 1;
    if (a.k & 1) {
      const uint32_t half = 1u << st;
      const int nbf = (R / 2) << a.m;
      for (int b = tid; b < nbf; b += nthreads) {
        uint32_t jj = b & (M - 1), bb = b >> a.m;
        uint32_t lo = bb & (half - 1);
        uint32_t t = ((bb >> st) << (st + 1)) | lo;
        int i0 = (t << a.m) + jj, i1 = ((t + half) << a.m) + jj;
        Fr x = lds_fr(s, i0), y = lds_fr(s, i1);
        Fr d = x - y;
        if (st > 0 && lo) d = d * lds_fr(itw, lo << (a.k - 1 - st));
        sts_fr(s, i0, x + y);
        sts_fr(s, i1, d);
      }
      __syncthreads();
      --st;
    }
    const int nb4 = (R / 4) << a.m;
    for (; st >= 1; st -= 2) {
      const uint32_t half = 1u << st, quarter = half >> 1;
      for (int b = tid; b < nb4; b += nthreads) {
        uint32_t jj = b & (M - 1), bb = b >> a.m;
        uint32_t lo = bb & (quarter - 1);
        uint32_t t0 = ((bb >> (st - 1)) << (st + 1)) | lo;
        int i0 = (t0 << a.m) + jj, i1 = ((t0 + quarter) << a.m) + jj, i2 = ((t0 + half) << a.m) + jj,
            i3 = ((t0 + half + quarter) << a.m) + jj;
        Fr x0 = lds_fr(s, i0), x1 = lds_fr(s, i1), x2 = lds_fr(s, i2), x3 = lds_fr(s, i3);
        const uint32_t e0 = lo << (a.k - 1 - st);
        Fr a0 = x0 + x2, a2 = x0 - x2, a1 = x1 + x3, a3 = x1 - x3;
        if (e0) a2 = a2 * lds_fr(itw, e0);
        a3 = a3 * lds_fr(itw, e0 + (R >> 2));
        Fr b0 = a0 + a1, b1 = a0 - a1, b2 = a2 + a3, b3 = a2 - a3;
        if (e0) { // level st-1: all twiddles are 1 on the last level and for lo == 0
          Fr w2 = lds_fr(itw, 2 * e0);
          b1 = b1 * w2;
          b3 = b3 * w2;
        }
        sts_fr(s, i0, b0);
        sts_fr(s, i1, b1);
        sts_fr(s, i2, b2);
        sts_fr(s, i3, b3);
      }
      __syncthreads();
    }

    for (int e = tid; e < R * M; e += nthreads) {
      uint32_t q, jj;
      if (a.ns_log == 0) { // first pass: the tile's outputs are one contiguous run, q fastest
        jj = e >> a.k;
        q = e & (R - 1);
      } else {
        q = e >> a.m;
        jj = e & (M - 1);
      }
      uint32_t j = j0 + jj;
      size_t idx = ((size_t)(j >> a.ns_log) << (a.ns_log + a.k)) + (j & ns_mask) + ((size_t)q << a.ns_log);
      uint32_t qr = a.k ? __brev(q) >> (32 - a.k) : 0;
      Fr v = lds_fr(s, (qr << a.m) + jj);
      if (a.last) {
        if (a.post_table)
          v = v * ld_fr(a.post_table + idx);
        else if (a.has_scale)
          v = v * a.scale;
      }
      st_fr(out + idx * a.estride, v);
    }
  }

  // 1-point "transform" and tiny helpers -------------------------------------------------------------
  static __global__ void __launch_bounds__(256)
    bitrev_swap_kernel(Fr* data, int logn, int batch, size_t batch_stride, int estride)
  {
    size_t n = (size_t)1 << logn, total = n * batch;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
      size_t b = g >> logn, i = g & (n - 1);
      size_t r = logn ? (size_t)(__brev((uint32_t)i) >> (32 - logn)) : 0;
      if (i < r) {
        Fr* p = data + b * batch_stride;
        Fr x = ld_fp_coherent(p + i * estride), y = ld_fp_coherent(p + r * estride);
        st_fr(p + i * estride, y);
        st_fr(p + r * estride, x);
      }
    }
  }

  // data[b][i] *= g^i  (g given as g^(2^j) table, Montgomery)
  static __global__ void __launch_bounds__(256)
    coset_mul_kernel(Fr* data, PowTable t, int logn, int batch, size_t batch_stride, int estride)
  {
    size_t n = (size_t)1 << logn, total = n * batch;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
      size_t b = g >> logn, i = g & (n - 1);
      Fr acc = Fr::one();
      for (int bit = 0; bit < logn; ++bit)
        if ((i >> bit) & 1) acc = acc * t.pw[bit];
      Fr* p = data + b * batch_stride + i * estride;
      st_fr(p, ld_fp_coherent(p) * acc);
    }
  }

  static void plan_passes(int logn, int* ks, int* npass)
  {
    const int kmax = ntt_tuning().kmax;
    int np = (logn + kmax - 1) / kmax;
    if (np < 1) np = 1;
    int base = logn / np, rem = logn % np;
    for (int i = 0; i < np; ++i)
      ks[i] = base + (i < rem ? 1 : 0);
    *npass = np;
  }

  eIcicleError ntt_enqueue(
    const Fr* in, Fr* out, int logn, bool inverse, int batch, bool columns_batch, const Fr* post_table, cudaStream_t st)
  {
    const NttDomain* dom = ntt_domain();
    if (!dom) return ICICLE_INVALID_ARGUMENT; // domain not initialised
    if (logn < 0 || logn > dom->max_log) return ICICLE_INVALID_ARGUMENT;
    if (batch < 1) return ICICLE_SUCCESS;
    const size_t n = (size_t)1 << logn;
    int ks[8], np;
    plan_passes(logn, ks, &np);
    if (logn == 0) ks[0] = 0;

    // ping-pong buffers so that the last pass lands in `out` and no pass reads what it writes
    const size_t bytes = n * batch * sizeof(Fr);
    Fr *s1 = nullptr, *s2 = nullptr;
    const bool inplace = (in == out);
    if (np > 1 || inplace) B200_CUDA(cudaMallocAsync((void**)&s1, bytes, st), ICICLE_ALLOCATION_FAILED);
    if (inplace && (np & 1) && np > 1) B200_CUDA(cudaMallocAsync((void**)&s2, bytes, st), ICICLE_ALLOCATION_FAILED);

    NttPassArgs a;
    a.tw = dom->tw;
    a.log_nmax = dom->max_log;
    a.logn = logn;
    a.estride = columns_batch ? batch : 1;
    a.batch_stride = columns_batch ? 1 : n;
    a.post_table = post_table;
    a.has_scale = inverse ? 1 : 0;
    if (inverse) {
      // 1/N = 2^-logn: invert on the host once (Fermat), Montgomery form
      Fr two = Fr::one().dbl(), acc = Fr::one();
      for (int i = 0; i < logn; ++i)
        acc = acc * two;
      a.scale = acc.inverse();
    } else {
      a.scale = Fr::one();
    }

    const Fr* src = in;
    int ns_log = 0;
    for (int p = 0; p < np; ++p) {
      Fr* dst;
      const int remaining = np - 1 - p; // passes after this one
      if (remaining == 0)
        dst = (inplace && np == 1) ? s1 : out;
      else if (s2)
        dst = (p == 0) ? s1 : (src == s1 ? s2 : s1); // in-place, odd count: in -> s1 -> s2 -> ... -> out
      else
        dst = (remaining & 1) ? s1 : out; // alternate so the final write is `out`
      if (s2 && remaining == 0) dst = out;
      a.in = src;
      a.out = dst;
      a.ns_log = ns_log;
      a.k = ks[p];
      int m = logn - ks[p];
      if (m > ntt_tuning().tile_log - ks[p]) m = ntt_tuning().tile_log - ks[p];
      if (ns_log > 0 && m > ns_log) m = ns_log;
      if (m < 0) m = 0;
      a.m = m;
      a.last = remaining == 0;
      const int tile = 1 << (a.k + a.m);
      const bool r4 = ntt_tuning().variant != 0;
      int threads = r4 ? tile / 4 : tile / 2;
      if (threads > (r4 ? 512 : 256)) threads = r4 ? 512 : 256;
      if (threads < 32) threads = 32;
      dim3 grid((unsigned)(n >> (a.k + a.m)), (unsigned)batch);
      size_t smem = ((size_t)tile + (size_t)(1 << a.k) / 2 + 1) * sizeof(Fr);
      if (smem > 48 * 1024) { // opt in to large dynamic shared memory once per kernel
        static std::once_flag once;
        std::call_once(once, [] {
          const int cap = 200 * 1024;
          cudaFuncSetAttribute(ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
          cudaFuncSetAttribute(ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
          cudaFuncSetAttribute(ntt_pass4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
          cudaFuncSetAttribute(ntt_pass4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        });
      }
      if (r4) {
        if (inverse)
          B200_LAUNCH(ntt_pass4_kernel<true>, grid, threads, smem, st, a);
        else
          B200_LAUNCH(ntt_pass4_kernel<false>, grid, threads, smem, st, a);
      } else if (inverse)
        B200_LAUNCH(ntt_pass_kernel<true>, grid, threads, smem, st, a);
      else
        B200_LAUNCH(ntt_pass_kernel<false>, grid, threads, smem, st, a);
      src = dst;
      ns_log += ks[p];
    }
    cudaError_t ce = cudaGetLastError();
    if (inplace && np == 1 && ce == cudaSuccess)
      ce = cudaMemcpyAsync(out, s1, bytes, cudaMemcpyDeviceToDevice, st);
    if (s1) cudaFreeAsync(s1, st);
    if (s2) cudaFreeAsync(s2, st);
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] ntt_enqueue: %s\n", cudaGetErrorString(ce));
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    return ICICLE_SUCCESS;
  }

} // namespace b200

using namespace b200;

extern "C" {

eIcicleError bn254_get_root_of_unity(uint64_t max_size, bn254_scalar_t* rou)
{
  if (!rou) return ICICLE_INVALID_POINTER;
  if (max_size == 0) return ICICLE_INVALID_ARGUMENT;
  int logn = 0;
  while (((uint64_t)1 << logn) < max_size && logn < 63)
    ++logn; // ceil(log2(max_size)) (icicle/src/ntt.cpp:56)
  if (logn > 28) return ICICLE_INVALID_ARGUMENT;
  Fr w = host_omega(logn);
  memcpy(rou, &w, 32);
  return ICICLE_SUCCESS;
}

eIcicleError bn254_ntt_init_domain(const bn254_scalar_t* primitive_root, const NTTInitDomainConfig* config)
{
  if (!primitive_root || !config) return ICICLE_INVALID_POINTER;
  Fr root;
  memcpy(&root, primitive_root, 32);
  return ntt_init_domain_host(root, as_stream(config->stream));
}

eIcicleError bn254_ntt_release_domain(void) { return ntt_release_domain_host(); }

eIcicleError bn254_get_root_of_unity_from_domain(uint64_t logn, bn254_scalar_t* rou)
{
  if (!rou) return ICICLE_INVALID_POINTER;
  const NttDomain* d = ntt_domain();
  if (!d || (int64_t)logn > d->max_log) return ICICLE_INVALID_ARGUMENT;
  Fr w = Fr::to_mont(d->root_std);
  for (int i = d->max_log; i > (int)logn; --i)
    w = w.sqr();
  w = Fr::from_mont(w);
  memcpy(rou, &w, 32);
  return ICICLE_SUCCESS;
}

eIcicleError
bn254_ntt(const bn254_scalar_t* input, int size, NTTDir dir, const NTTConfig* cfg, bn254_scalar_t* output)
{
  if (!input || !output || !cfg) return ICICLE_INVALID_POINTER;
  if (size <= 0 || (size & (size - 1))) return ICICLE_INVALID_ARGUMENT; // the reference throws (ntt.cuh:676-682)
  B200_TRY(ensure_device());
  const NttDomain* dom = ntt_domain();
  if (!dom) return ICICLE_INVALID_ARGUMENT;
  int logn = 0;
  while ((1 << logn) < size)
    ++logn;
  if (logn > dom->max_log) return ICICLE_INVALID_ARGUMENT; // reference throws (ntt.cuh:669-674)
  const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  const bool inverse = dir == kInverse;
  cudaStream_t st = as_stream(cfg->stream);
  const size_t total = (size_t)size * batch, bytes = total * sizeof(Fr);

  StagedIn I;
  StagedOut O;
  B200_TRY(I.init(input, bytes, cfg->are_inputs_on_device, st));
  B200_TRY(O.init(output, bytes, cfg->are_outputs_on_device, st));
  const Fr* src = (const Fr*)I.dev;
  Fr* dst = (Fr*)O.dev;

  const int estride = cfg->columns_batch ? batch : 1;
  const size_t bstride = cfg->columns_batch ? 1 : (size_t)size;
  const int ord = cfg->ordering;
  const bool rev_in = (ord == kRN || ord == kRR), rev_out = (ord == kNR || ord == kRR);

  // coset: forward multiplies the input by g^j, inverse multiplies the output by g^-j
  // (mixed_radix_ntt.cu:976-983, 1009-1014). coset_gen == 0 or 1 means "no coset".
  Fr g;
  memcpy(&g, &cfg->coset_gen, 32);
  Fr raw1 = Fr::raw_one();
  const bool coset = !g.is_zero() && g != raw1;
  PowTable gt;
  if (coset) {
    Fr gm = Fr::to_mont(g);
    if (inverse) gm = gm.inverse();
    for (int i = 0; i < 28; ++i) {
      gt.pw[i] = gm;
      gm = gm.sqr();
    }
  }

  eIcicleError err = ICICLE_SUCCESS;
  cudaError_t ce = cudaSuccess;
  const bool pre_touch = rev_in || (coset && !inverse);
  if (pre_touch) {
    // work on a private copy in `dst` so the caller's input is never modified
    if (src != dst) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
    src = dst;
    if (rev_in) B200_LAUNCH(bitrev_swap_kernel, grid_for(total, 256, 8), 256, 0, st, dst, logn, batch, bstride, estride);
    if (coset && !inverse)
      B200_LAUNCH(coset_mul_kernel, grid_for(total, 256, 8), 256, 0, st, dst, gt, logn, batch, bstride, estride);
  }
  if (ce == cudaSuccess) err = ntt_enqueue(src, dst, logn, inverse, batch, cfg->columns_batch, nullptr, st);
  if (err == ICICLE_SUCCESS && ce == cudaSuccess) {
    if (coset && inverse)
      B200_LAUNCH(coset_mul_kernel, grid_for(total, 256, 8), 256, 0, st, dst, gt, logn, batch, bstride, estride);
    if (rev_out) B200_LAUNCH(bitrev_swap_kernel, grid_for(total, 256, 8), 256, 0, st, dst, logn, batch, bstride, estride);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) err = translate(ce, ICICLE_UNKNOWN_FALLBACK);
  if (err == ICICLE_SUCCESS) err = O.finish(st);
  I.release(st);
  if (err != ICICLE_SUCCESS) return err;
  if (!cfg->is_async) B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED);
  return ICICLE_SUCCESS;
}

} // extern "C"
