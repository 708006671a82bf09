// Kernels + host orchestration of the Pippenger MSM (see msm.cuh for the design table).
// Included by msm_g1.cu (F = Fq) and msm_g2.cu (F = Fq2) so the two instantiations compile in parallel.
#pragma once
#include "msm.cuh"

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  static constexpr uint32_t DIGIT_NONE = 0xffffffffu;
  static constexpr int REDUCE_CHUNK = 16; // buckets per running-sum thread

  struct MsmDev { // plan fields the kernels need, passed by value
    int n, c, windows, factor, sets, bpw, nbuckets, item_cap;
    uint32_t h[9];
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

  // ------------------------------------------------------------------------------------------------
  // (1) signed-digit decomposition + bucket histogram.
  // s' = s + H, raw window u_w = bits [cw, cw+c) of s', digit d_w = u_w - 2^(c-1) in [-2^(c-1), 2^(c-1)).
  // digits[w*n + i] = (|d|-1) | sign<<31, or DIGIT_NONE for d == 0.  Bucket key = (w % sets)*bpw + |d|-1.
  // Replaces split_scalars_kernel (cuda_msm.cuh:166-203).
  static __global__ void __launch_bounds__(256)
    msm_digits_kernel(MsmDev pl, const Fr* scalars, bool scalars_mont, uint32_t* digits, uint32_t* hist)
  {
    const uint32_t half = 1u << (pl.c - 1);
    const uint32_t mask = (1u << pl.c) - 1;
    const int lane = threadIdx.x & 31;
    const int n_round = (pl.n + 31) & ~31; // whole warps stay in the loop so the warp-wide match below is convergent
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
      const bool live = i < pl.n;
      Fr s = live ? ld_fr(scalars + i) : Fr::zero();
      if (scalars_mont) s = Fr::from_mont(s);
      uint32_t t[10];
      uint64_t carry = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        carry += (uint64_t)s.v[k] + pl.h[k];
        t[k] = (uint32_t)carry;
        carry >>= 32;
      }
      t[8] = (uint32_t)carry + pl.h[8];
      t[9] = 0;
      for (int w = 0; w < pl.windows; ++w) {
        int bit = w * pl.c;
        int limb = bit >> 5, sh = bit & 31;
        uint64_t two = ((uint64_t)t[limb + 1] << 32) | t[limb];
        uint32_t u = (uint32_t)(two >> sh) & mask;
        uint32_t out = DIGIT_NONE, key = DIGIT_NONE;
        if (live && u != half) {
          uint32_t neg = u < half;
          uint32_t mag = neg ? half - u : u - half; // 1..half
          out = (mag - 1) | (neg << 31);
          key = (w % pl.sets) * pl.bpw + (mag - 1);
        }
        // warp-aggregated histogram: skewed scalars (0/1-heavy witnesses) and short top windows put
        // millions of entries on a handful of keys; one atomic per distinct key per warp
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (key != DIGIT_NONE && lane == __ffs(peers) - 1) atomicAdd(&hist[key], (uint32_t)__popc(peers));
        if (live) digits[(size_t)w * pl.n + i] = out;
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // (2) exclusive scan (three small kernels; inputs are a few MB and L2-resident)
  static constexpr int SCAN_BLOCK = 1024, SCAN_PER_THREAD = 4, SCAN_TILE = SCAN_BLOCK * SCAN_PER_THREAD;

  static __global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_kernel(const uint32_t* in, int n, uint32_t* out, uint32_t* tile_sums)
  {
    __shared__ uint32_t warp_sums[32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
      v[k] = (base + k < n) ? in[base + k] : 0;
      sum += v[k];
    }
    uint32_t incl = sum;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_sums[lane] = wi - ws; // exclusive
      if (lane == 31) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
      if (base + k < n) out[base + k] = run;
      run += v[k];
    }
  }

  // single CTA: exclusive scan of tile sums in place; writes the grand total to *total
  static __global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* tile_sums, int ntiles, uint32_t* total)
  {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
      int i = base + threadIdx.x;
      uint32_t v = i < ntiles ? tile_sums[i] : 0, incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) warp_sums[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
          if (lane >= d) wi += o;
        }
        warp_sums[lane] = wi - ws;
      }
      __syncthreads();
      uint32_t excl = carry_s + warp_sums[wid] + incl - v;
      if (i < ntiles) tile_sums[i] = excl;
      __syncthreads();
      if (threadIdx.x == 1023) carry_s = excl + v;
      __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
  }

  static __global__ void __launch_bounds__(SCAN_BLOCK) scan_add_kernel(uint32_t* out, int n, const uint32_t* tile_sums, const uint32_t* total)
  {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k)
      if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total; // out has n+1 slots
  }

  // out[0..n] = exclusive scan of in[0..n), out[n] = total. tile_sums: >= ceil(n/SCAN_TILE)+1 words.
  static inline void exclusive_scan(const uint32_t* in, int n, uint32_t* out, uint32_t* tile_sums, cudaStream_t st)
  {
    int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    B200_LAUNCH(scan_tile_kernel, ntiles, SCAN_BLOCK, 0, st, in, n, out, tile_sums);
    B200_LAUNCH(scan_sums_kernel, 1, 1024, 0, st, tile_sums, ntiles, tile_sums + ntiles);
    B200_LAUNCH(scan_add_kernel, ntiles, SCAN_BLOCK, 0, st, out, n, tile_sums, tile_sums + ntiles);
  }

  // ------------------------------------------------------------------------------------------------
  // (3) scatter point references into bucket order, window-major so the write window stays in L2.
  // entries[pos] = (i*f + w/sets) | sign<<31.  cursor starts as a copy of the exclusive offsets.
  static __global__ void __launch_bounds__(256)
    msm_scatter_kernel(MsmDev pl, const uint32_t* digits, uint32_t* cursor, uint32_t* entries)
  {
    const size_t total = (size_t)pl.n * pl.windows;
    const size_t total_round = (total + 31) & ~(size_t)31;
    const int lane = threadIdx.x & 31;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total_round; e += (size_t)gridDim.x * blockDim.x) {
      uint32_t d = e < total ? digits[e] : DIGIT_NONE;
      int w = (int)(e / pl.n), i = (int)(e - (size_t)w * pl.n);
      uint32_t key = d == DIGIT_NONE ? DIGIT_NONE : (w % pl.sets) * pl.bpw + (d & 0x7fffffffu);
      // warp-aggregated cursor bump: the leader of each group of equal keys reserves the whole run
      uint32_t peers = __match_any_sync(0xffffffffu, key);
      int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (key != DIGIT_NONE && lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (key != DIGIT_NONE) {
        uint32_t pos = base + __popc(peers & ((1u << lane) - 1));
        entries[pos] = (uint32_t)(i * pl.factor + w / pl.sets) | (d & 0x80000000u);
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // (4) work items: bucket k with n_k entries becomes ceil(n_k/T) items of <= T entries.
  struct MsmItem {
    uint32_t begin;  // first entry
    uint32_t len;    // 1..T
    uint32_t bucket; // bucket key
    uint32_t dst;    // index into `buckets` (single-item bucket) or 0x80000000|index into `partials`
  };

  static __global__ void __launch_bounds__(256) msm_item_count_kernel(MsmDev pl, const uint32_t* offsets, uint32_t* nitems)
  {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < pl.nbuckets; k += gridDim.x * blockDim.x) {
      uint32_t cnt = offsets[k + 1] - offsets[k];
      nitems[k] = (cnt + pl.item_cap - 1) / pl.item_cap;
    }
  }

  // writes the items of every bucket (unsorted, bucket-major) + the histogram of item lengths;
  // empty buckets are cleared here; buckets with >1 item are appended to the `multi` list.
  template <class F>
  __global__ void __launch_bounds__(256) msm_item_build_kernel(
    MsmDev pl, const uint32_t* offsets, const uint32_t* item_off, MsmItem* items, uint32_t* len_hist, XYZZ<F>* buckets,
    uint32_t* multi, uint32_t* multi_count)
  {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < pl.nbuckets; k += gridDim.x * blockDim.x) {
      uint32_t beg = offsets[k], cnt = offsets[k + 1] - beg;
      uint32_t first = item_off[k], ni = item_off[k + 1] - first;
      if (ni == 0) {
        st_struct(buckets + k, XYZZ<F>::inf());
        continue;
      }
      if (ni > 1) multi[atomicAdd(multi_count, 1u)] = (uint32_t)k;
      for (uint32_t j = 0; j < ni; ++j) {
        uint32_t len = min((uint32_t)pl.item_cap, cnt - j * pl.item_cap);
        MsmItem it;
        it.begin = beg + j * pl.item_cap;
        it.len = len;
        it.bucket = (uint32_t)k;
        it.dst = ni == 1 ? (uint32_t)k : (0x80000000u | (first + j));
        reinterpret_cast<uint4*>(items)[first + j] = make_uint4(it.begin, it.len, it.bucket, it.dst);
        atomicAdd(&len_hist[pl.item_cap - len], 1u); // descending-length order: bin 0 = longest
      }
    }
  }

  static __global__ void __launch_bounds__(256)
    msm_item_sort_kernel(const MsmItem* items, const uint32_t* total_items, int item_cap, uint32_t* len_cursor, MsmItem* sorted)
  {
    uint32_t n = *total_items;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
      uint4 it = reinterpret_cast<const uint4*>(items)[t];
      uint32_t pos = atomicAdd(&len_cursor[item_cap - it.y], 1u);
      reinterpret_cast<uint4*>(sorted)[pos] = it;
    }
  }

  // ------------------------------------------------------------------------------------------------
  // (5) bucket accumulation: one thread per work item, serial mixed adds over its entries with the
  // next point prefetched into registers while the current add runs. Hot loop #1
  // (replaces accumulate_buckets_kernel, cuda_msm.cuh:223-255).
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
    const MsmItem* sorted, const uint32_t* total_items, const uint32_t* entries, const Affine<F>* bases, XYZZ<F>* buckets,
    XYZZ<F>* partials)
  {
    uint32_t n = *total_items;
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
    const uint32_t* multi, const uint32_t* multi_count, const uint32_t* item_off, const XYZZ<F>* partials, XYZZ<F>* buckets)
  {
    uint32_t nm = *multi_count;
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
    const uint32_t* multi, const uint32_t* multi_count, const uint32_t* item_off, const XYZZ<F>* partials, XYZZ<F>* buckets)
  {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    uint32_t nm = *multi_count;
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
  // (6) bucket reduction.  Per set: sum_b (b+1) * B_b.  Thread (set, chunk q) runs the classic running
  // sum over REDUCE_CHUNK buckets (tot = sum (j+1) B_j, run = sum B_j) and adds (q*CHUNK)*run by a short
  // double-and-add.  Replaces the log-halving passes of cuda_msm.cuh:846-942.
  template <class F>
  __global__ void __launch_bounds__(128) msm_reduce_chunks_kernel(MsmDev pl, const XYZZ<F>* buckets, XYZZ<F>* chunk_sums)
  {
    int chunks_per_set = pl.bpw / REDUCE_CHUNK;
    if (chunks_per_set == 0) chunks_per_set = 1;
    int chunk_len = pl.bpw < REDUCE_CHUNK ? pl.bpw : REDUCE_CHUNK;
    int total = pl.sets * chunks_per_set;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      int set = t / chunks_per_set, q = t - set * chunks_per_set;
      const XYZZ<F>* b = buckets + (size_t)set * pl.bpw + (size_t)q * chunk_len;
      XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
      for (int j = chunk_len - 1; j >= 0; --j) {
        xyzz_add_ni(run, ld_struct(b + j));
        xyzz_add_ni(tot, run);
      }
      uint32_t base = (uint32_t)q * chunk_len;
      if (base) {
        XYZZ<F> m = XYZZ<F>::inf();
        for (int bit = 31 - __clz(base); bit >= 0; --bit) {
          xyzz_dbl_ni(m);
          if ((base >> bit) & 1) xyzz_add_ni(m, run);
        }
        xyzz_add_ni(tot, m);
      }
      st_struct(chunk_sums + t, tot);
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

  // Horner over sets (weights 2^(c*set)); writes the reference's boundary layout
  template <class F>
  __global__ void msm_final_kernel(MsmDev pl, const XYZZ<F>* set_sums, Projective<F>* out_std)
  {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = ld_struct(set_sums + pl.sets - 1);
    for (int s = pl.sets - 2; s >= 0; --s) {
      for (int k = 0; k < pl.c; ++k)
        xyzz_dbl_ni(acc);
      xyzz_add_ni(acc, ld_struct(set_sums + s));
    }
    Projective<F> p = acc.to_projective();
    Projective<F> o = {F::from_mont(p.x), F::from_mont(p.y), F::from_mont(p.z)};
    st_struct(out_std, o);
  }

  // ------------------------------------------------------------------------------------------------
  template <class F>
  __global__ void __launch_bounds__(128)
    msm_precompute_kernel(const Affine<F>* in, bool in_mont, int n, int factor, int shift, Affine<F>* out, bool out_mont)
  {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      Affine<F> a = ld_affine(in + i);
      if (!in_mont) a = {F::to_mont(a.x), F::to_mont(a.y)};
      XYZZ<F> p = XYZZ<F>::from_affine(a);
      for (int j = 0; j < factor; ++j) {
        if (j) {
          for (int k = 0; k < shift; ++k)
            xyzz_dbl_ni(p);
          a = p.to_affine();
          p = XYZZ<F>::from_affine(a); // keep ZZ = 1 so later doublings stay cheap to normalise
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

  // ------------------------------------------------------------------------------------------------
  template <class F>
  eIcicleError msm_enqueue(
    const MsmPlan& plan, const Fr* scalars, bool scalars_mont, const Affine<F>* bases, Projective<F>* out_std, cudaStream_t st)
  {
    MsmDev pl;
    pl.n = plan.n;
    pl.c = plan.c;
    pl.windows = plan.windows;
    pl.factor = plan.factor;
    pl.sets = plan.sets;
    pl.bpw = plan.bpw;
    pl.nbuckets = plan.nbuckets;
    pl.item_cap = plan.item_cap;
    for (int i = 0; i < 9; ++i)
      pl.h[i] = plan.hconst[i];

    const size_t ne = plan.entries();
    const int nb = plan.nbuckets;
    const size_t max_items = (size_t)nb + ne / plan.item_cap + 1;
    const int scan_tiles_b = (nb + SCAN_TILE - 1) / SCAN_TILE + 2;
    const int chunks_per_set = plan.bpw / REDUCE_CHUNK > 0 ? plan.bpw / REDUCE_CHUNK : 1;

    // one stream-ordered scratch block, carved up (pool keeps it across calls)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_digits = 0;
    size_t o_entries = o_digits + al(ne * 4);
    size_t o_hist = o_entries + al(ne * 4);
    size_t o_offsets = o_hist + al((size_t)nb * 4);
    size_t o_cursor = o_offsets + al((size_t)(nb + 1) * 4);
    size_t o_nitems = o_cursor + al((size_t)(nb + 1) * 4);
    size_t o_itemoff = o_nitems + al((size_t)nb * 4);
    size_t o_tiles = o_itemoff + al((size_t)(nb + 1) * 4);
    size_t o_lenhist = o_tiles + al((size_t)scan_tiles_b * 4);
    size_t o_lenoff = o_lenhist + al((size_t)(plan.item_cap + 1) * 4);
    size_t o_small = o_lenoff + al((size_t)(plan.item_cap + 2) * 4); // [0]=multi_count
    size_t o_multi = o_small + 256;
    size_t o_items = o_multi + al((size_t)nb * 4);
    size_t o_sorted = o_items + al(max_items * sizeof(MsmItem));
    size_t o_buckets = o_sorted + al(max_items * sizeof(MsmItem));
    size_t o_partials = o_buckets + al((size_t)nb * sizeof(XYZZ<F>));
    size_t o_chunks = o_partials + al(max_items * sizeof(XYZZ<F>));
    size_t o_sets = o_chunks + al((size_t)plan.sets * chunks_per_set * sizeof(XYZZ<F>));
    size_t total = o_sets + al((size_t)plan.sets * sizeof(XYZZ<F>));

    uint8_t* base = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&base, total, st), ICICLE_ALLOCATION_FAILED);
    uint32_t* digits = (uint32_t*)(base + o_digits);
    uint32_t* entries = (uint32_t*)(base + o_entries);
    uint32_t* hist = (uint32_t*)(base + o_hist);
    uint32_t* offsets = (uint32_t*)(base + o_offsets);
    uint32_t* cursor = (uint32_t*)(base + o_cursor);
    uint32_t* nitems = (uint32_t*)(base + o_nitems);
    uint32_t* item_off = (uint32_t*)(base + o_itemoff);
    uint32_t* tiles = (uint32_t*)(base + o_tiles);
    uint32_t* len_hist = (uint32_t*)(base + o_lenhist);
    uint32_t* len_off = (uint32_t*)(base + o_lenoff);
    uint32_t* multi_count = (uint32_t*)(base + o_small);
    uint32_t* multi = (uint32_t*)(base + o_multi);
    MsmItem* items = (MsmItem*)(base + o_items);
    MsmItem* sorted = (MsmItem*)(base + o_sorted);
    XYZZ<F>* buckets = (XYZZ<F>*)(base + o_buckets);
    XYZZ<F>* partials = (XYZZ<F>*)(base + o_partials);
    XYZZ<F>* chunk_sums = (XYZZ<F>*)(base + o_chunks);
    XYZZ<F>* set_sums = (XYZZ<F>*)(base + o_sets);

    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t e) {
      if (ce == cudaSuccess) ce = e;
    };
    // hist .. small are contiguous: one memset clears hist, len_hist and multi_count (others are overwritten)
    chk(cudaMemsetAsync(hist, 0, o_offsets - o_hist, st));
    chk(cudaMemsetAsync(len_hist, 0, o_multi - o_lenhist, st));

    const int sms = sm_count();
    B200_LAUNCH(msm_digits_kernel, grid_for(plan.n, 256, 8), 256, 0, st, pl, scalars, scalars_mont, digits, hist);
    exclusive_scan(hist, nb, offsets, tiles, st);
    chk(cudaMemcpyAsync(cursor, offsets, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st));
    B200_LAUNCH(msm_scatter_kernel, grid_for(ne, 256, 8), 256, 0, st, pl, digits, cursor, entries);
    B200_LAUNCH(msm_item_count_kernel, grid_for(nb, 256, 8), 256, 0, st, pl, offsets, nitems);
    exclusive_scan(nitems, nb, item_off, tiles, st); // item_off[nb] = total items (device-side only)
    B200_LAUNCH(
      msm_item_build_kernel<F>, grid_for(nb, 256, 8), 256, 0, st, pl, offsets, item_off, items, len_hist, buckets, multi,
      multi_count);
    exclusive_scan(len_hist, plan.item_cap + 1, len_off, tiles, st);
    B200_LAUNCH(msm_item_sort_kernel, grid_for(max_items, 256, 8), 256, 0, st, items, item_off + nb, plan.item_cap, len_off, sorted);
    if (g_profile_events[0]) cudaEventRecord(g_profile_events[0], st);
    B200_LAUNCH(
      msm_accumulate_kernel<F>, grid_for(max_items, 128, 16), 128, 0, st, sorted, item_off + nb, entries, bases, buckets,
      partials);
    if (g_profile_events[1]) cudaEventRecord(g_profile_events[1], st);
    B200_LAUNCH(msm_fold_serial_kernel<F>, grid_for(nb, 128, 8), 128, 0, st, multi, multi_count, item_off, partials, buckets);
    B200_LAUNCH(
      msm_fold_kernel<F>, sms, FOLD_BLOCK, FOLD_BLOCK * sizeof(XYZZ<F>), st, multi, multi_count, item_off, partials, buckets);
    B200_LAUNCH(
      msm_reduce_chunks_kernel<F>, grid_for((size_t)plan.sets * chunks_per_set, 128, 16), 128, 0, st, pl, buckets, chunk_sums);
    {
      // level 1: G CTAs per set, level 2: one CTA per set over the G partial sums
      int G = (chunks_per_set + WSUM_BLOCK * 4 - 1) / (WSUM_BLOCK * 4);
      if (G > 128) G = 128;
      if (G < 1) G = 1;
      int per_cta = (chunks_per_set + G - 1) / G;
      const size_t sm = WSUM_BLOCK * sizeof(XYZZ<F>);
      if (G == 1) {
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(1, plan.sets), WSUM_BLOCK, sm, st, chunk_sums, chunks_per_set, chunks_per_set, set_sums);
      } else {
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(G, plan.sets), WSUM_BLOCK, sm, st, chunk_sums, chunks_per_set, per_cta, partials);
        B200_LAUNCH(msm_set_sum_kernel<F>, dim3(1, plan.sets), WSUM_BLOCK, sm, st, partials, G, G, set_sums);
      }
    }
    B200_LAUNCH(msm_final_kernel<F>, 1, 32, 0, st, pl, set_sums, out_std);
    chk(cudaGetLastError());
    cudaFreeAsync(base, st);
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] msm_enqueue: %s\n", cudaGetErrorString(ce));
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    return ICICLE_SUCCESS;
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
    if ((size_t)msm_size * f >= (1ull << 31)) return ICICLE_INVALID_ARGUMENT;

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
      MsmPlan plan = make_msm_plan(msm_size, cfg->c, bitsize, f, g2);
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
