// Kernels + host orchestration of the Pippenger MSM (see msm.cuh for the design table).
// Included by msm_g1.cu (F = Fq) and msm_g2.cu (F = Fq2) so the two instantiations compile in parallel.
#pragma once
#include "msm.cuh"
#include "field_inv.cuh"

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  static constexpr int REDUCE_CHUNK = 8; // buckets per running-sum thread
  static constexpr int MSM_MAX_SEL = 4;   // base-point sets one sort can feed in a single accumulate launch

  template <class F>
  struct BasesSel { // the base-point arrays of the MSMs that share one sort (blockIdx.y selects)
    const Affine<F>* p[MSM_MAX_SEL];
  };

  template <class F>
  __device__ __forceinline__ Affine<F> ld_affine(const Affine<F>* p)
  {
    constexpr int NQ = sizeof(Affine<F>) / 16;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 t[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      t[i] = __ldg(q + i);
    Affine<F> r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      w[4 * i] = t[i].x;
      w[4 * i + 1] = t[i].y;
      w[4 * i + 2] = t[i].z;
      w[4 * i + 3] = t[i].w;
    }
    return r;
  }

  template <class T>
  __device__ __forceinline__ T ld_struct(const T* p) // coherent 128-bit loads of a 16 B-aligned struct
  {
    constexpr int NQ = sizeof(T) / 16;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    T r;
    uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      w[i] = q[i];
    return r;
  }
  template <class T>
  __device__ __forceinline__ void st_struct(T* p, const T& v)
  {
    constexpr int NQ = sizeof(T) / 16;
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* w = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      q[i] = w[i];
  }


  // Cold kernels (reduction, folds, precompute) call the group law out of line: keeps their code size and
  // the build time down; only the accumulate loop inlines it.
  template <class F>
  __device__ __noinline__ void xyzz_add_ni(XYZZ<F>& a, const XYZZ<F>& b)
  {
    a.add(b);
  }
  template <class F>
  __device__ __noinline__ void xyzz_dbl_ni(XYZZ<F>& a)
  {
    a = a.dbl();
  }

  template <class F>
  struct AccTraits { // G1: 128 registers, next point prefetched into registers
    static constexpr bool kPrefetch = true;
    static constexpr int kMinBlocks = 4;
  };
  template <>
  struct AccTraits<Fq2> { // G2: the accumulator alone is 64 registers; no register prefetch, 3 CTAs/SM instead
    static constexpr bool kPrefetch = false;
    static constexpr int kMinBlocks = 3;
  };

  template <class F>
  __global__ void __launch_bounds__(128, AccTraits<F>::kMinBlocks) msm_accumulate_kernel(
    const MsmItem* sorted, const uint32_t* total_items, const uint32_t* entries, BasesSel<F> sel, XYZZ<F>* buckets,
    XYZZ<F>* partials, uint32_t nbuckets, uint32_t max_items)
  {
    uint32_t n = *total_items;
    const Affine<F>* bases = sel.p[blockIdx.y]; // MSMs sharing this sort: one grid row each
    buckets += (size_t)blockIdx.y * nbuckets;
    partials += (size_t)blockIdx.y * max_items;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
      uint4 it = reinterpret_cast<const uint4*>(sorted)[t];
      const uint32_t* e = entries + it.x;
      uint32_t len = it.y;
      XYZZ<F> acc = XYZZ<F>::inf();
      if (AccTraits<F>::kPrefetch) {
        uint32_t cur = e[0];
        Affine<F> p = ld_affine(bases + (cur & 0x7fffffffu));
        for (uint32_t k = 0; k < len; ++k) {
          uint32_t nxt = cur;
          Affine<F> pn = p;
          if (k + 1 < len) {
            nxt = e[k + 1];
            pn = ld_affine(bases + (nxt & 0x7fffffffu));
          }
          if (cur >> 31) p.y = p.y.neg();
          acc.madd(p);
          cur = nxt;
          p = pn;
        }
      } else {
        uint32_t cur = e[0];
        for (uint32_t k = 0; k < len; ++k) {
          Affine<F> p = ld_affine(bases + (cur & 0x7fffffffu));
          uint32_t sign = cur >> 31;
          if (k + 1 < len) cur = e[k + 1];
          if (sign) p.y = p.y.neg();
          acc.madd(p);
        }
      }
      XYZZ<F>* dst = (it.w >> 31) ? partials + (it.w & 0x7fffffffu) : buckets + it.w;
      st_struct(dst, acc);
    }
  }

  // block-wide sum of `count` XYZZ values at src[0..count); result valid in thread 0
  template <class F, int BLOCK>
  __device__ __forceinline__ XYZZ<F> block_sum(const XYZZ<F>* src, uint32_t count, XYZZ<F>* sh)
  {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = threadIdx.x; i < count; i += BLOCK)
      xyzz_add_ni(acc, ld_struct(src + i));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        XYZZ<F> a = sh[threadIdx.x];
        xyzz_add_ni(a, sh[threadIdx.x + s]);
        sh[threadIdx.x] = a;
      }
      __syncthreads();
    }
    return sh[0];
  }

  static constexpr int FOLD_BLOCK = 64;
  static constexpr uint32_t FOLD_SERIAL_MAX = 32; // buckets cut into <= this many items are folded by one thread

  // (5b) buckets that were cut into several items. Moderately long buckets (the common case when one bucket
  // set serves all windows) are folded by one thread each; giant ones (skewed scalars) by one CTA each.
  template <class F>
  __global__ void __launch_bounds__(128) msm_fold_serial_kernel(
    const uint32_t* multi, const uint32_t* multi_count, const uint32_t* item_off, const XYZZ<F>* partials, XYZZ<F>* buckets,
    uint32_t nbuckets, uint32_t max_items)
  {
    uint32_t nm = *multi_count;
    buckets += (size_t)blockIdx.y * nbuckets;
    partials += (size_t)blockIdx.y * max_items;
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < nm; m += gridDim.x * blockDim.x) {
      uint32_t k = multi[m];
      uint32_t first = item_off[k], ni = item_off[k + 1] - first;
      if (ni > FOLD_SERIAL_MAX) continue;
      XYZZ<F> acc = ld_struct(partials + first);
      for (uint32_t j = 1; j < ni; ++j)
        xyzz_add_ni(acc, ld_struct(partials + first + j));
      st_struct(buckets + k, acc);
    }
  }

  template <class F>
  __global__ void __launch_bounds__(FOLD_BLOCK) msm_fold_kernel(
    const uint32_t* multi, const uint32_t* multi_count, const uint32_t* item_off, const XYZZ<F>* partials, XYZZ<F>* buckets,
    uint32_t nbuckets, uint32_t max_items)
  {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    uint32_t nm = *multi_count;
    buckets += (size_t)blockIdx.y * nbuckets;
    partials += (size_t)blockIdx.y * max_items;
    for (uint32_t m = blockIdx.x; m < nm; m += gridDim.x) {
      uint32_t k = multi[m];
      uint32_t first = item_off[k], ni = item_off[k + 1] - first;
      if (ni <= FOLD_SERIAL_MAX) continue; // uniform per CTA
      XYZZ<F> r = block_sum<F, FOLD_BLOCK>(partials + first, ni, sh);
      if (threadIdx.x == 0) st_struct(buckets + k, r);
      __syncthreads();
    }
  }

  // ------------------------------------------------------------------------------------------------
  // (6) bucket reduction.  Per set: W = sum_b (b+1) * B_b, as a two-level running sum with no scalar multiplication
  // on the wide level.  Level 0: thread (set, chunk q) runs the classic running sum over L0 = REDUCE_CHUNK buckets and
  // emits T0_q = sum_j (j+1) B_{qL0+j} and R0_q = sum_j B_{qL0+j}; then W = sum_q T0_q + L0 * sum_q q R0_q.
  // Level 1 (1/L0 of the elements): the same running sum over L2 consecutive R0 values gives T1, R1 and
  // sum_q q R0_q = sum_q2 [T1_q2 + (L2 q2 - 1) R1_q2]; only this level pays a short double-and-add.  Its outputs,
  // pre-multiplied by L0, are stored behind the T0 values of the set so that one tree sum finishes the set.
  // Replaces the log-halving passes of cuda_msm.cuh:846-942.
  template <class F>
  __global__ void __launch_bounds__(128) msm_reduce_chunks_kernel(
    MsmDev pl, int nsel, const uint32_t* offsets, const XYZZ<F>* buckets, XYZZ<F>* chunk_sums, XYZZ<F>* chunk_runs, int out_stride)
  {
    int chunks_per_set = pl.bpw / REDUCE_CHUNK;
    if (chunks_per_set == 0) chunks_per_set = 1;
    int chunk_len = pl.bpw < REDUCE_CHUNK ? pl.bpw : REDUCE_CHUNK;
    int per_msm = pl.sets * chunks_per_set;
    int total = nsel * per_msm;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      int which = t / per_msm, r = t - which * per_msm;
      int set = r / chunks_per_set, q = r - set * chunks_per_set;
      size_t key0 = (size_t)set * pl.bpw + (size_t)q * chunk_len;
      const XYZZ<F>* b = buckets + (size_t)which * pl.nbuckets + key0;
      const uint32_t* off = offsets + key0;
      XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
      for (int j = chunk_len - 1; j >= 0; --j) {
        if (off[j + 1] != off[j]) xyzz_add_ni(run, ld_struct(b + j)); // empty buckets were never written
        xyzz_add_ni(tot, run);
      }
      st_struct(chunk_sums + (size_t)(which * pl.sets + set) * out_stride + q, tot);
      if (chunks_per_set > 1) st_struct(chunk_runs + t, run);
    }
  }

  // level 1 of the bucket reduction: thread (set, q2) over L2 consecutive level-0 chunk sums (see above)
  template <class F>
  __global__ void __launch_bounds__(128) msm_reduce_level1_kernel(
    int nsets, int chunks_per_set, int l0_log, const XYZZ<F>* chunk_runs, XYZZ<F>* chunk_sums, int out_stride)
  {
    const int l2 = chunks_per_set < REDUCE_CHUNK ? chunks_per_set : REDUCE_CHUNK;
    const int chunks2 = chunks_per_set / l2;
    const int total = nsets * chunks2;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      int set = t / chunks2, q2 = t - set * chunks2;
      const XYZZ<F>* r = chunk_runs + (size_t)set * chunks_per_set + (size_t)q2 * l2;
      XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
      for (int j = l2 - 1; j >= 0; --j) {
        xyzz_add_ni(run, ld_struct(r + j));
        xyzz_add_ni(tot, run);
      }
      // tot = sum (j+1) R0_j ; wanted: sum (q2 l2 + j) R0_j = tot + (q2 l2 - 1) run
      uint32_t base = (uint32_t)q2 * l2;
      if (base) {
        XYZZ<F> m = XYZZ<F>::inf();
        for (int bit = 31 - __clz(base); bit >= 0; --bit) {
          xyzz_dbl_ni(m);
          if ((base >> bit) & 1) xyzz_add_ni(m, run);
        }
        xyzz_add_ni(tot, m);
      }
      xyzz_add_ni(tot, run.neg());
      for (int k = 0; k < l0_log; ++k)
        xyzz_dbl_ni(tot);
      st_struct(chunk_sums + (size_t)set * out_stride + chunks_per_set + q2, tot);
    }
  }

  static constexpr int WSUM_BLOCK = 128;

  // sums `per_cta` consecutive values per CTA: in[set][g*per_cta ..] -> out[set*gridDim.x + g]; two levels of this
  // reduce the chunk sums of a set to one value without a single long serial loop
  template <class F>
  __global__ void __launch_bounds__(WSUM_BLOCK)
    msm_set_sum_kernel(const XYZZ<F>* in, int per_set, int per_cta, XYZZ<F>* out)
  {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    int beg = blockIdx.x * per_cta;
    int cnt = per_set - beg < per_cta ? per_set - beg : per_cta;
    if (cnt < 0) cnt = 0;
    XYZZ<F> r = block_sum<F, WSUM_BLOCK>(in + (size_t)blockIdx.y * per_set + beg, (uint32_t)cnt, sh);
    if (threadIdx.x == 0) st_struct(out + (size_t)blockIdx.y * gridDim.x + blockIdx.x, r);
  }

  // Horner over sets (weights 2^(c*set)); writes the reference's boundary layout. One block per MSM of the selection.
  template <class F>
  __global__ void msm_final_kernel(MsmDev pl, const XYZZ<F>* set_sums, Projective<F>* out_std)
  {
    if (threadIdx.x != 0) return;
    set_sums += (size_t)blockIdx.x * pl.sets;
    XYZZ<F> acc = ld_struct(set_sums + pl.sets - 1);
    for (int s = pl.sets - 2; s >= 0; --s) {
      for (int k = 0; k < pl.c; ++k)
        xyzz_dbl_ni(acc);
      xyzz_add_ni(acc, ld_struct(set_sums + s));
    }
    Projective<F> p = acc.to_projective();
    Projective<F> o = {F::from_mont(p.x), F::from_mont(p.y), F::from_mont(p.z)};
    st_struct(out_std + blockIdx.x, o);
  }

  // ------------------------------------------------------------------------------------------------
  // x^-1 through the division-step inversion (field_inv.cuh: ~57 products' worth, mostly on the ALU pipe) instead of
  // Fermat's ~380 multiplier-bound products; Fq2 by the norm
  template <class Cfg>
  __device__ __forceinline__ Fp<Cfg> inverse_fast(const Fp<Cfg>& x)
  {
    return inverse_safegcd(x);
  }
  __device__ __forceinline__ Fq2 inverse_fast(const Fq2& x)
  {
    Fq n = inverse_safegcd(x.c0.sqr() + x.c1.sqr());
    return {x.c0 * n, (x.c1 * n).neg()};
  }

  // Table j of point i is 2^(shift*j) P_i: `shift` Jacobian doublings (a = 0: dbl-2009-l, 2M + 5S - the XYZZ doubling the
  // buckets use costs 9) from the previous, normalised entry, then one inversion per entry.  The cold path of the prover
  // (cache build) is dominated by this kernel: 12 extra tables x 17 M points at 3200k constraints.
  template <class F>
  __global__ void __launch_bounds__(128)
    msm_precompute_kernel(const Affine<F>* in, bool in_mont, int n, int factor, int shift, Affine<F>* out, bool out_mont)
  {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      Affine<F> a = ld_affine(in + i);
      if (!in_mont) a = {F::to_mont(a.x), F::to_mont(a.y)};
      for (int j = 0; j < factor; ++j) {
        if (j && !a.is_inf()) {
          F X = a.x, Y = a.y, Z = F::one();
#pragma unroll 1
          for (int k = 0; k < shift; ++k) {
            F A = X.sqr(), B = Y.sqr(), C = B.sqr();
            F D = ((X + B).sqr() - A - C).dbl();
            F E = A.dbl() + A;
            Z = (Y * Z).dbl();
            X = E.sqr() - D.dbl();
            Y = E * (D - X) - C.dbl().dbl().dbl();
          }
          // Z = 0 only for a point of order two (none on BN254; arbitrary caller data): inverse(0) = 0 -> (0, 0) = infinity
          F zi = inverse_fast(Z), zi2 = zi.sqr();
          a = {X * zi2, Y * (zi2 * zi)};
        }
        Affine<F> o = out_mont ? a : Affine<F>{F::from_mont(a.x), F::from_mont(a.y)};
        st_struct(out + (size_t)i * factor + j, o);
      }
    }
  }

  template <class F>
  eIcicleError precompute_enqueue(
    const Affine<F>* in, bool in_mont, int n, int factor, int shift, Affine<F>* out, bool out_mont, cudaStream_t st)
  {
    if (n <= 0) return ICICLE_SUCCESS;
    B200_LAUNCH(msm_precompute_kernel<F>, grid_for(n, 128), 128, 0, st, in, in_mont, n, factor, shift, out, out_mont);
    B200_CUDA(cudaGetLastError(), ICICLE_UNKNOWN_FALLBACK);
    return ICICLE_SUCCESS;
  }

} // namespace b200
#include "msm_batch_affine.cuh"
#include "msm_reduce_quad.cuh"
namespace b200 {

  // ---- quad-cooperative variants of the reduction kernels (msm_reduce_quad.cuh): one quad per running-sum chunk ----
  inline bool reduce_quad_enabled()
  {
    static const bool on = [] {
      const char* e = getenv("B200_MSM_QUAD");
      return !(e && e[0] == '0');
    }();
    return on;
  }

  template <class F>
  __global__ void __launch_bounds__(128) msm_reduce_chunks_quad_kernel(
    MsmDev pl, int nsel, const uint32_t* offsets, const XYZZ<F>* buckets, XYZZ<F>* chunk_sums, XYZZ<F>* chunk_runs, int out_stride)
  {
    int chunks_per_set = pl.bpw / REDUCE_CHUNK;
    if (chunks_per_set == 0) chunks_per_set = 1;
    int chunk_len = pl.bpw < REDUCE_CHUNK ? pl.bpw : REDUCE_CHUNK;
    int per_msm = pl.sets * chunks_per_set;
    int total = nsel * per_msm;
    const int ql = threadIdx.x & 3;
    const unsigned qm = 0xFu << (threadIdx.x & 28);
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; t < total; t += (gridDim.x * blockDim.x) >> 2) {
      int which = t / per_msm, r = t - which * per_msm;
      int set = r / chunks_per_set, q = r - set * chunks_per_set;
      size_t key0 = (size_t)set * pl.bpw + (size_t)q * chunk_len;
      const XYZZ<F>* b = buckets + (size_t)which * pl.nbuckets + key0;
      const uint32_t* off = offsets + key0;
      XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
      for (int j = chunk_len - 1; j >= 0; --j) {
        if (off[j + 1] != off[j]) xyzz_add_quad(run, ld_struct(b + j), ql, qm); // empty buckets were never written
        xyzz_add_quad(tot, run, ql, qm);
      }
      if (ql == 0) {
        st_struct(chunk_sums + (size_t)(which * pl.sets + set) * out_stride + q, tot);
        if (chunks_per_set > 1) st_struct(chunk_runs + t, run);
      }
    }
  }

  template <class F>
  __global__ void __launch_bounds__(128) msm_reduce_level1_quad_kernel(
    int nsets, int chunks_per_set, int l0_log, const XYZZ<F>* chunk_runs, XYZZ<F>* chunk_sums, int out_stride)
  {
    const int l2 = chunks_per_set < REDUCE_CHUNK ? chunks_per_set : REDUCE_CHUNK;
    const int chunks2 = chunks_per_set / l2;
    const int total = nsets * chunks2;
    const int ql = threadIdx.x & 3;
    const unsigned qm = 0xFu << (threadIdx.x & 28);
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; t < total; t += (gridDim.x * blockDim.x) >> 2) {
      int set = t / chunks2, q2 = t - set * chunks2;
      const XYZZ<F>* r = chunk_runs + (size_t)set * chunks_per_set + (size_t)q2 * l2;
      XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
      for (int j = l2 - 1; j >= 0; --j) {
        xyzz_add_quad(run, ld_struct(r + j), ql, qm);
        xyzz_add_quad(tot, run, ql, qm);
      }
      // tot = sum (j+1) R0_j ; wanted: sum (q2 l2 + j) R0_j = tot + (q2 l2 - 1) run
      uint32_t base = (uint32_t)q2 * l2;
      if (base) {
        XYZZ<F> m = XYZZ<F>::inf();
        for (int bit = 31 - __clz(base); bit >= 0; --bit) {
          xyzz_dbl_quad(m, ql, qm);
          if ((base >> bit) & 1) xyzz_add_quad(m, run, ql, qm);
        }
        xyzz_add_quad(tot, m, ql, qm);
      }
      xyzz_add_quad(tot, run.neg(), ql, qm);
      for (int k = 0; k < l0_log; ++k)
        xyzz_dbl_quad(tot, ql, qm);
      if (ql == 0) st_struct(chunk_sums + (size_t)set * out_stride + chunks_per_set + q2, tot);
    }
  }

  // sums `per_cta` consecutive values per CTA like msm_set_sum_kernel: the 32 quads of the CTA take strided values, then a
  // tree over the quads' partial sums in shared memory
  template <class F>
  __global__ void __launch_bounds__(WSUM_BLOCK)
    msm_set_sum_quad_kernel(const XYZZ<F>* in, int per_set, int per_cta, XYZZ<F>* out)
  {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    constexpr int NQ = WSUM_BLOCK / 4;
    const int ql = threadIdx.x & 3, qi = threadIdx.x >> 2;
    const unsigned qm = 0xFu << (threadIdx.x & 28);
    int beg = blockIdx.x * per_cta;
    int cnt = per_set - beg < per_cta ? per_set - beg : per_cta;
    if (cnt < 0) cnt = 0;
    const XYZZ<F>* src = in + (size_t)blockIdx.y * per_set + beg;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = qi; i < cnt; i += NQ)
      xyzz_add_quad(acc, ld_struct(src + i), ql, qm);
    if (ql == 0) sh[qi] = acc;
    __syncthreads();
    for (int s = NQ / 2; s > 0; s >>= 1) {
      if (qi < s) {
        XYZZ<F> a = sh[qi];
        xyzz_add_quad(a, sh[qi + s], ql, qm);
        if (ql == 0) sh[qi] = a;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) st_struct(out + (size_t)blockIdx.y * gridDim.x + blockIdx.x, sh[0]);
  }


  // ------------------------------------------------------------------------------------------------
  template <class F>
  eIcicleError msm_reduce_enqueue(
    const MsmPlan& plan, const MsmSorted& sorted, const Affine<F>* const* bases_mont, int nsel, Projective<F>* out_std,
    cudaStream_t st, cudaEvent_t gate)
  {
    if (nsel < 1 || nsel > MSM_MAX_SEL) return ICICLE_INVALID_ARGUMENT;
    MsmDev pl = msm_dev_plan(plan);
    const int nb = plan.nbuckets;
    const size_t max_items = sorted.max_items;
    const int chunks_per_set = plan.bpw / REDUCE_CHUNK > 0 ? plan.bpw / REDUCE_CHUNK : 1;
    const int nsets = plan.sets * nsel;
    // level-1 outputs of the bucket reduction live behind the level-0 sums of each set
    const int l2 = chunks_per_set < REDUCE_CHUNK ? chunks_per_set : REDUCE_CHUNK;
    const int chunks2 = chunks_per_set > 1 ? chunks_per_set / l2 : 0;
    const int sum_stride = chunks_per_set + chunks2;
    int l0_log = 0;
    while ((1 << l0_log) < (plan.bpw < REDUCE_CHUNK ? plan.bpw : REDUCE_CHUNK)) ++l0_log;

    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_buckets = 0;
    size_t o_partials = o_buckets + al((size_t)nsel * nb * sizeof(XYZZ<F>));
    size_t o_chunks = o_partials + al((size_t)nsel * max_items * sizeof(XYZZ<F>));
    size_t o_runs = o_chunks + al((size_t)nsets * sum_stride * sizeof(XYZZ<F>));
    size_t o_sets = o_runs + al((size_t)nsets * chunks_per_set * sizeof(XYZZ<F>));
    size_t total = o_sets + al((size_t)nsets * sizeof(XYZZ<F>));
    uint8_t* base = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&base, total, st), ICICLE_ALLOCATION_FAILED);
    XYZZ<F>* buckets = (XYZZ<F>*)(base + o_buckets);
    XYZZ<F>* partials = (XYZZ<F>*)(base + o_partials);
    XYZZ<F>* chunk_sums = (XYZZ<F>*)(base + o_chunks);
    XYZZ<F>* chunk_runs = (XYZZ<F>*)(base + o_runs);
    XYZZ<F>* set_sums = (XYZZ<F>*)(base + o_sets);

    BasesSel<F> sel;
    for (int k = 0; k < MSM_MAX_SEL; ++k)
      sel.p[k] = bases_mont[k < nsel ? k : 0];
    const uint32_t* total_items = sorted.item_off + nb;
    const int sms = sm_count();
    const unsigned ysel = (unsigned)nsel;

    const int ba_rounds = msm_batch_affine_rounds(plan, sizeof(F) > sizeof(Fq));
    if (gate) cudaStreamWaitEvent(st, gate, 0);
    msm_profile_begin(st);
    if (ba_rounds) {
      // long buckets: pairwise tree of batched affine additions (msm_batch_affine.cuh), 6-7 products per add
      eIcicleError be = msm_accumulate_batched_enqueue<F>(plan, sorted, sel, nsel, ba_rounds, buckets, st);
      if (be != ICICLE_SUCCESS) {
        cudaFreeAsync(base, st);
        return be;
      }
    } else {
      B200_LAUNCH(
        msm_accumulate_kernel<F>, dim3(grid_for(max_items, 128, 16), ysel), 128, 0, st, sorted.sorted, total_items, sorted.entries,
        sel, buckets, partials, (uint32_t)nb, (uint32_t)max_items);
      B200_LAUNCH(
        msm_fold_serial_kernel<F>, dim3(grid_for(nb, 128, 8), ysel), 128, 0, st, sorted.multi, sorted.multi_count, sorted.item_off,
        partials, buckets, (uint32_t)nb, (uint32_t)max_items);
      B200_LAUNCH(
        msm_fold_kernel<F>, dim3(sms, ysel), FOLD_BLOCK, FOLD_BLOCK * sizeof(XYZZ<F>), st, sorted.multi, sorted.multi_count,
        sorted.item_off, partials, buckets, (uint32_t)nb, (uint32_t)max_items);
    }
    msm_profile_end(st, plan, sizeof(F) > sizeof(Fq) ? 1 : 0, nsel, ba_rounds);
    // Level 0 is multiplier work when there are many chunks (one thread each); with few it is a latency chain like the
    // levels above it, which always run quad-cooperatively (msm_reduce_quad.cuh)
    const bool quad = reduce_quad_enabled();
    if (quad && (size_t)nsets * chunks_per_set <= 32768)
      B200_LAUNCH(
        msm_reduce_chunks_quad_kernel<F>, grid_for((size_t)nsets * chunks_per_set * 4, 128, 16), 128, 0, st, pl, nsel, sorted.offsets,
        buckets, chunk_sums, chunk_runs, sum_stride);
    else
      B200_LAUNCH(
        msm_reduce_chunks_kernel<F>, grid_for((size_t)nsets * chunks_per_set, 128, 16), 128, 0, st, pl, nsel, sorted.offsets, buckets,
        chunk_sums, chunk_runs, sum_stride);
    if (chunks2 > 0) {
      if (quad)
        B200_LAUNCH(
          msm_reduce_level1_quad_kernel<F>, grid_for((size_t)nsets * chunks2 * 4, 128, 16), 128, 0, st, nsets, chunks_per_set, l0_log,
          chunk_runs, chunk_sums, sum_stride);
      else
        B200_LAUNCH(
          msm_reduce_level1_kernel<F>, grid_for((size_t)nsets * chunks2, 128, 16), 128, 0, st, nsets, chunks_per_set, l0_log, chunk_runs,
          chunk_sums, sum_stride);
    }
    {
      // level 1: G CTAs per set, level 2: one CTA per set over the G partial sums
      int G = (sum_stride + WSUM_BLOCK * 4 - 1) / (WSUM_BLOCK * 4);
      if (G > 128) G = 128;
      if (G < 1) G = 1;
      int per_cta = (sum_stride + G - 1) / G;
      const size_t sm = WSUM_BLOCK * sizeof(XYZZ<F>);
      if (quad) {
        if (G == 1) {
          B200_LAUNCH(msm_set_sum_quad_kernel<F>, dim3(1, nsets), WSUM_BLOCK, sm, st, chunk_sums, sum_stride, sum_stride, set_sums);
        } else {
          B200_LAUNCH(msm_set_sum_quad_kernel<F>, dim3(G, nsets), WSUM_BLOCK, sm, st, chunk_sums, sum_stride, per_cta, partials);
          B200_LAUNCH(msm_set_sum_quad_kernel<F>, dim3(1, nsets), WSUM_BLOCK, sm, st, partials, G, G, set_sums);
        }
      } else if (G == 1) {
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(1, nsets), WSUM_BLOCK, sm, st, chunk_sums, sum_stride, sum_stride, set_sums);
      } else {
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(G, nsets), WSUM_BLOCK, sm, st, chunk_sums, sum_stride, per_cta, partials);
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(1, nsets), WSUM_BLOCK, sm, st, partials, G, G, set_sums);
      }
    }
    B200_LAUNCH(msm_final_kernel<F>, nsel, 32, 0, st, pl, set_sums, out_std);
    cudaError_t ce = cudaGetLastError();
    cudaFreeAsync(base, st);
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] msm_reduce_enqueue: %s\n", cudaGetErrorString(ce));
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    return ICICLE_SUCCESS;
  }

  template <class F>
  eIcicleError msm_enqueue(
    const MsmPlan& plan, const Fr* scalars, bool scalars_mont, const Affine<F>* bases, Projective<F>* out_std, cudaStream_t st)
  {
    MsmSorted sorted;
    B200_TRY(msm_sort_enqueue(plan, scalars, scalars_mont, &sorted, st));
    eIcicleError e = msm_reduce_enqueue<F>(plan, sorted, &bases, 1, out_std, st);
    msm_sorted_free(&sorted, st);
    return e;
  }

  // ------------------------------------------------------------------------------------------------
  // C-ABI body shared by bn254_msm / bn254_g2_msm (icicle/src/msm.cpp:12-32 -> cuda_msm.cuh:1397-1443)
  template <class F>
  __global__ void __launch_bounds__(256) affine_to_mont_kernel(const Affine<F>* in, size_t n, Affine<F>* out)
  {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      Affine<F> a = ld_affine(in + i);
      Affine<F> o = {F::to_mont(a.x), F::to_mont(a.y)};
      st_struct(out + i, o);
    }
  }

  // sum of k partial results (boundary layout: homogeneous projective, standard form) -> out; one thread
  template <class F>
  __global__ void msm_combine_chunks_kernel(const Projective<F>* parts, int k, Projective<F>* out)
  {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = 0; i < k; ++i) {
      Projective<F> p = ld_struct(parts + i);
      Projective<F> m = {F::to_mont(p.x), F::to_mont(p.y), F::to_mont(p.z)};
      xyzz_add_ni(acc, xyzz_from_projective(m));
    }
    Projective<F> r = acc.to_projective();
    Projective<F> o = {F::from_mont(r.x), F::from_mont(r.y), F::from_mont(r.z)};
    st_struct(out, o);
  }

  // Largest number of points one Pippenger pass may take: 32-bit entry positions (n * windows < 2^32), 31-bit table
  // indices (n * f < 2^31), and the device memory the pass needs (staging of host operands, the sort scratch, the
  // Montgomery copy of the points) within 70 % of what is free now.  B200_MSM_CHUNK=<points> forces a smaller chunk
  // (tests).  The reference splits for the same reasons (cuda_msm.cuh:1130-1237 multi-chunk, :1239-1394 on memory).
  inline size_t msm_chunk_points(const MsmPlan& full, int f, size_t affine_bytes, bool host_scalars, bool host_points, bool mont_copy)
  {
    size_t lim = ((1ull << 32) - 1) / (size_t)full.windows;
    const size_t lim_f = ((1ull << 31) - 1) / (size_t)f;
    if (lim_f < lim) lim = lim_f;
    size_t per_point = (size_t)full.windows * 4 * 2                    // digits + entries
                       + (size_t)full.windows * 2 * sizeof(MsmItem) / (size_t)full.item_cap + 1;
    if (host_scalars) per_point += 2 * 32;                             // double-buffered staging
    if (host_points) per_point += 2 * affine_bytes * (size_t)f;
    if (mont_copy) per_point += affine_bytes * (size_t)f;
    // device memory: only problems that could matter are worth a driver query (cudaMemGetInfo is not free, and small MSMs
    // are latency-sensitive): below 8 GiB of estimated need the answer cannot change the plan on a 180 GB part
    const size_t fixed = (size_t)full.nbuckets * (4 * 8 + 2 * sizeof(MsmItem) + 3 * 2 * affine_bytes); // offsets, items, buckets
    if (fixed + per_point * (size_t)full.n > ((size_t)8 << 30)) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const size_t budget = (size_t)(0.7 * (double)free_b);
        const size_t by_mem = budget > fixed ? (budget - fixed) / per_point : 1;
        if (by_mem < lim) lim = by_mem;
      }
    }
    if (const char* e = getenv("B200_MSM_CHUNK")) {
      const long long v = atoll(e);
      if (v > 0 && (size_t)v < lim) lim = (size_t)v;
    }
    return lim < 1 ? 1 : lim;
  }

  // One MSM larger than a single pass may take (msm_chunk_points) or with host-resident operands that should not be
  // staged whole: K passes over contiguous chunks, host operands double-buffered on a copy stream so chunk k+1 crosses
  // PCIe while chunk k is computed, partial sums added by one small kernel.  Every chunk uses the full problem's window
  // width so a precomputed table (built for that width) stays valid.
  template <class F>
  eIcicleError msm_chunked(
    const MsmPlan& full, const Fr* scalars, bool scalars_dev, const Affine<F>* bases, bool bases_dev, size_t n, size_t chunk,
    const MSMConfig* cfg, int bitsize, bool g2, Projective<F>* out_dev, cudaStream_t st)
  {
    const int f = full.stride;
    const int K = (int)((n + chunk - 1) / chunk);
    const bool mont_copy = !cfg->are_points_montgomery_form;
    cudaStream_t cs = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    Fr* sbuf[2] = {nullptr, nullptr};
    Affine<F>* pbuf[2] = {nullptr, nullptr};
    Affine<F>* mont_tmp = nullptr;
    Projective<F>* parts = nullptr;
    eIcicleError err = ICICLE_SUCCESS;
    auto ck = [&](cudaError_t e, eIcicleError code) {
      if (e != cudaSuccess && err == ICICLE_SUCCESS) {
        fprintf(stderr, "[icicle_b200] chunked msm: %s\n", cudaGetErrorString(e));
        err = code;
      }
      return e == cudaSuccess;
    };
    const bool staged = !scalars_dev || !bases_dev;
    if (staged) {
      ck(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), ICICLE_STREAM_CREATION_FAILED);
      for (int i = 0; i < 2 && err == ICICLE_SUCCESS; ++i) {
        ck(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming), ICICLE_UNKNOWN_FALLBACK);
        ck(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming), ICICLE_UNKNOWN_FALLBACK);
        if (!scalars_dev) ck(cudaMallocAsync((void**)&sbuf[i], chunk * sizeof(Fr), st), ICICLE_ALLOCATION_FAILED);
        if (!bases_dev) ck(cudaMallocAsync((void**)&pbuf[i], chunk * f * sizeof(Affine<F>), st), ICICLE_ALLOCATION_FAILED);
      }
      if (err == ICICLE_SUCCESS) { // the copy stream may touch the buffers once they exist in st's order
        ck(cudaEventRecord(consumed[0], st), ICICLE_UNKNOWN_FALLBACK);
        ck(cudaEventRecord(consumed[1], st), ICICLE_UNKNOWN_FALLBACK);
      }
    }
    if (err == ICICLE_SUCCESS && mont_copy) ck(cudaMallocAsync((void**)&mont_tmp, chunk * f * sizeof(Affine<F>), st), ICICLE_ALLOCATION_FAILED);
    if (err == ICICLE_SUCCESS) ck(cudaMallocAsync((void**)&parts, (size_t)K * sizeof(Projective<F>), st), ICICLE_ALLOCATION_FAILED);
    auto stage = [&](int k) { // enqueue the host->device copies of chunk k on the copy stream
      const size_t lo = (size_t)k * chunk, cn = std::min(chunk, n - lo);
      const int b = k & 1;
      ck(cudaStreamWaitEvent(cs, consumed[b], 0), ICICLE_UNKNOWN_FALLBACK);
      if (!scalars_dev) ck(cudaMemcpyAsync(sbuf[b], scalars + lo, cn * sizeof(Fr), cudaMemcpyHostToDevice, cs), ICICLE_COPY_FAILED);
      if (!bases_dev)
        ck(cudaMemcpyAsync(pbuf[b], bases + lo * f, cn * f * sizeof(Affine<F>), cudaMemcpyHostToDevice, cs), ICICLE_COPY_FAILED);
      ck(cudaEventRecord(copied[b], cs), ICICLE_UNKNOWN_FALLBACK);
    };
    if (err == ICICLE_SUCCESS && staged) stage(0);
    for (int k = 0; k < K && err == ICICLE_SUCCESS; ++k) {
      const size_t lo = (size_t)k * chunk, cn = std::min(chunk, n - lo);
      const int b = k & 1;
      if (staged) {
        if (k + 1 < K) stage(k + 1);
        ck(cudaStreamWaitEvent(st, copied[b], 0), ICICLE_UNKNOWN_FALLBACK);
      }
      const Fr* sc = scalars_dev ? scalars + lo : sbuf[b];
      const Affine<F>* pts = bases_dev ? bases + lo * f : pbuf[b];
      if (mont_copy && err == ICICLE_SUCCESS) {
        B200_LAUNCH(affine_to_mont_kernel<F>, grid_for(cn * f, 256, 8), 256, 0, st, pts, cn * f, mont_tmp);
        pts = mont_tmp;
      }
      MsmPlan plan = make_msm_plan((int)cn, full.c, bitsize, f, g2);
      if (err == ICICLE_SUCCESS) err = msm_enqueue<F>(plan, sc, cfg->are_scalars_montgomery_form, pts, parts + k, st);
      if (staged) ck(cudaEventRecord(consumed[b], st), ICICLE_UNKNOWN_FALLBACK);
    }
    if (err == ICICLE_SUCCESS) {
      B200_LAUNCH(msm_combine_chunks_kernel<F>, 1, 32, 0, st, parts, K, out_dev);
      ck(cudaGetLastError(), ICICLE_UNKNOWN_FALLBACK);
    }
    for (int i = 0; i < 2; ++i) {
      if (sbuf[i]) cudaFreeAsync(sbuf[i], st);
      if (pbuf[i]) cudaFreeAsync(pbuf[i], st);
    }
    if (mont_tmp) cudaFreeAsync(mont_tmp, st);
    if (parts) cudaFreeAsync(parts, st);
    if (staged) {
      // the copy stream and its events must outlive the work that references them
      if (err != ICICLE_SUCCESS) cudaStreamSynchronize(st);
      if (cs) {
        cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
      }
      for (int i = 0; i < 2; ++i) {
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed[i]) cudaEventDestroy(consumed[i]);
      }
    }
    return err;
  }

  template <class F>
  eIcicleError msm_api(const void* scalars_v, const void* bases_v, int msm_size, const MSMConfig* cfg, void* results_v, bool g2)
  {
    if (!cfg || !scalars_v || !bases_v || !results_v) return ICICLE_INVALID_POINTER;
    if (msm_size < 0) return ICICLE_INVALID_ARGUMENT;
    B200_TRY(ensure_device());
    cudaStream_t st = as_stream(cfg->stream);
    const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
    const int f = cfg->precompute_factor > 1 ? cfg->precompute_factor : 1;
    const int bitsize = cfg->bitsize > 0 ? cfg->bitsize : 254;
    const bool shared = cfg->are_points_shared_in_batch || batch == 1;
    const bool s_dev = is_device_ptr(scalars_v, cfg->are_scalars_on_device);
    const bool p_dev = is_device_ptr(bases_v, cfg->are_points_on_device);

    // one pass when it fits (32-bit entry positions, table indices, device memory); otherwise chunks
    MsmPlan plan = make_msm_plan(msm_size > 0 ? msm_size : 1, cfg->c, bitsize, f, g2);
    const size_t chunk = msm_size > 0 ? msm_chunk_points(plan, f, sizeof(Affine<F>), !s_dev, !p_dev, !cfg->are_points_montgomery_form) : 1;
    if (msm_size > 0 && chunk < (size_t)msm_size) {
      StagedOut O;
      B200_TRY(O.init(results_v, (size_t)batch * sizeof(Projective<F>), cfg->are_results_on_device, st));
      eIcicleError err = ICICLE_SUCCESS;
      for (int b = 0; b < batch && err == ICICLE_SUCCESS; ++b)
        err = msm_chunked<F>(
          plan, (const Fr*)scalars_v + (size_t)b * msm_size, s_dev,
          (const Affine<F>*)bases_v + (shared ? 0 : (size_t)b * msm_size * f), p_dev, (size_t)msm_size, chunk, cfg, bitsize, g2,
          (Projective<F>*)O.dev + b, st);
      if (err == ICICLE_SUCCESS) err = O.finish(st);
      if (err != ICICLE_SUCCESS) return err;
      if (!cfg->is_async) B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED);
      return ICICLE_SUCCESS;
    }

    const size_t n_scalars = (size_t)msm_size * batch;
    const size_t n_points = (size_t)msm_size * f * (shared ? 1 : batch);
    StagedIn S, P;
    StagedOut O;
    B200_TRY(S.init(scalars_v, n_scalars * sizeof(Fr), cfg->are_scalars_on_device, st));
    B200_TRY(P.init(bases_v, n_points * sizeof(Affine<F>), cfg->are_points_on_device, st));
    B200_TRY(O.init(results_v, (size_t)batch * sizeof(Projective<F>), cfg->are_results_on_device, st));

    const Affine<F>* bases = (const Affine<F>*)P.dev;
    Affine<F>* mont_tmp = nullptr;
    if (!cfg->are_points_montgomery_form && n_points) {
      B200_CUDA(cudaMallocAsync((void**)&mont_tmp, n_points * sizeof(Affine<F>), st), ICICLE_ALLOCATION_FAILED);
      B200_LAUNCH(affine_to_mont_kernel<F>, grid_for(n_points, 256, 8), 256, 0, st, bases, n_points, mont_tmp);
      bases = mont_tmp;
    }
    eIcicleError err = ICICLE_SUCCESS;
    if (msm_size == 0) {
      // empty sum = identity (0,1,0)
      Projective<F> id = {F::zero(), F::from_mont(F::one()), F::zero()};
      for (int b = 0; b < batch && err == ICICLE_SUCCESS; ++b)
        if (cudaMemcpyAsync((Projective<F>*)O.dev + b, &id, sizeof(id), cudaMemcpyHostToDevice, st) != cudaSuccess)
          err = ICICLE_COPY_FAILED;
      cudaStreamSynchronize(st); // `id` is a stack temporary
    } else {
      for (int b = 0; b < batch && err == ICICLE_SUCCESS; ++b) {
        err = msm_enqueue<F>(
          plan, (const Fr*)S.dev + (size_t)b * msm_size, cfg->are_scalars_montgomery_form,
          bases + (shared ? 0 : (size_t)b * msm_size * f), (Projective<F>*)O.dev + b, st);
      }
    }
    if (mont_tmp) cudaFreeAsync(mont_tmp, st);
    if (err == ICICLE_SUCCESS) err = O.finish(st);
    S.release(st);
    P.release(st);
    if (err != ICICLE_SUCCESS) return err;
    if (!cfg->is_async) B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED);
    return ICICLE_SUCCESS;
  }

  template <class F>
  eIcicleError precompute_api(const void* in_v, int n, const MSMConfig* cfg, void* out_v, bool g2)
  {
    if (!cfg || !in_v || !out_v) return ICICLE_INVALID_POINTER;
    if (n < 0) return ICICLE_INVALID_ARGUMENT;
    B200_TRY(ensure_device());
    cudaStream_t st = as_stream(cfg->stream);
    const int f = cfg->precompute_factor > 1 ? cfg->precompute_factor : 1;
    const int batch = (cfg->are_points_shared_in_batch || cfg->batch_size < 1) ? 1 : cfg->batch_size;
    const size_t total = (size_t)n * batch;
    if (total * f >= (1ull << 31)) return ICICLE_INVALID_ARGUMENT;
    StagedIn P;
    StagedOut O;
    B200_TRY(P.init(in_v, total * sizeof(Affine<F>), cfg->are_points_on_device, st));
    // the reference writes the table to host or device according to are_results_on_device (cuda_msm.cuh:1477-1530)
    B200_TRY(O.init(out_v, total * f * sizeof(Affine<F>), cfg->are_results_on_device, st));
    // same c as the msm call will pick for this (n, f): shift = c * sets (cuda_msm.cuh:1465)
    MsmPlan plan = make_msm_plan(n > 0 ? n : 1, cfg->c, cfg->bitsize > 0 ? cfg->bitsize : 254, f, g2);
    eIcicleError err = precompute_enqueue<F>(
      (const Affine<F>*)P.dev, cfg->are_points_montgomery_form, (int)total, f, plan.c * plan.sets, (Affine<F>*)O.dev,
      cfg->are_points_montgomery_form, st);
    if (err == ICICLE_SUCCESS) err = O.finish(st);
    P.release(st);
    if (err != ICICLE_SUCCESS) return err;
    if (!cfg->is_async) B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED);
    return ICICLE_SUCCESS;
  }

} // namespace b200
