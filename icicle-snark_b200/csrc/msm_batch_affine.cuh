// Bucket accumulation by batched affine addition.  Replaces msm_accumulate_kernel + the fold
// kernels for one group of MSMs sharing a sort when the buckets are long enough (msm_batch_affine_rounds()).
//
// The reference sums a bucket serially in projective coordinates (cuda_msm.cuh:223-255, 12M+ per add); the XYZZ
// kernel of msm_impl.cuh does it in 10 products.  Here every bucket is reduced as a pairwise tree of AFFINE additions,
// 6 products each once the denominators' inverses are known (batch_affine.cuh):
//
//   round r:  slot s of the output = in[src(s)] + in[src(s)+1]   (or a copied odd tail), len_next = ceil(len / 2)
//
// One kernel per round.  A CTA of 128 threads owns a tile of 128*K output slots; thread t owns slots tile + i*128 + t
// (i < K), so that every step of the tile touches 128 consecutive slots.  Forward pass: the running products of the K
// denominators go to a scratch array (coalesced, 32 B per slot); one inversion per thread by division steps
// (field_inv.cuh: ~57 products' worth of mostly ALU work instead of Fermat's ~380 multiplier-bound ones); backward pass:
// the operands are re-read and the K additions finished.  K is chosen per round (<= 64) so that the tiles fill whole
// waves of resident CTAs.  After the rounds each bucket holds <= a few points, summed by one thread (XYZZ), long
// leftovers (skewed scalars would need more rounds) by one CTA each.  Slot descriptors (operand indices) are produced
// per round by a bucket-parallel kernel from the segment offsets; a warp shares the work of a long bucket.
// Variants that were built and measured, then dropped (profiles/r02_batch_affine.md): prefix products in shared memory
// with one inverting warp per CTA (barrier-bound: the inversion's latency exceeds a tile's multiplier time), cp.async
// staging of the next slot's operands (slower on the gather round), a start stagger between co-resident CTAs.
#pragma once
#include <cstdlib>
#include <cstring>

#include "batch_affine.cuh"

namespace b200 {

  template <class F>
  struct BaCfg { // G1: up to 64 slots per thread and inversion, 5 CTAs/SM (<= 102 registers)
    static constexpr int K = 64;
    static constexpr int kMinBlocks = 5;
  };
  template <>
  struct BaCfg<Fq2> { // G2: the operands alone are 128 registers
    static constexpr int K = 32;
    static constexpr int kMinBlocks = 3;
  };
  static constexpr int BA_BLOCK = 128;
  static constexpr uint32_t BA_FIN_SERIAL = 48; // leftover segments longer than this are summed by a CTA

  // len_next[b] = ceil(len[b] / 2)
  static __global__ void __launch_bounds__(256) ba_halve_kernel(const uint32_t* off, int nb, uint32_t* len_next)
  {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
      uint32_t L = off[b + 1] - off[b];
      len_next[b] = (L + 1) >> 1;
    }
  }

  static constexpr uint32_t BA_NONE = 0xffffffffu;

  // Slot descriptors: desc[off_next[b] + j] = {first operand, second operand or BA_NONE (odd tail, copied)} for
  // j < ceil(len / 2).  Round 0: the operands are the sorted entry words themselves (point index | sign << 31), later
  // rounds: indices into the previous round's output.  Lanes own buckets; a warp shares the work of a long one.
  static __global__ void __launch_bounds__(256)
    ba_desc_kernel(const uint32_t* off, const uint32_t* off_next, int nb, const uint32_t* entries, uint2* desc)
  {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; b0 < nb; b0 += warps * 32) {
      const int b = b0 + lane;
      uint32_t src = 0, L = 0, o = 0;
      if (b < nb) {
        src = off[b];
        L = off[b + 1] - src;
        o = off_next[b];
      }
      const uint32_t slots = (L + 1) >> 1;
      const bool is_long = slots > 32;
      if (!is_long)
        for (uint32_t j = 0; j < slots; ++j) {
          const uint32_t k = src + 2 * j;
          const bool pair = 2 * j + 1 < L;
          desc[o + j] = entries ? make_uint2(entries[k], pair ? entries[k + 1] : BA_NONE) : make_uint2(k, pair ? k + 1 : BA_NONE);
        }
      uint32_t m = __ballot_sync(0xffffffffu, is_long);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t s2 = __shfl_sync(0xffffffffu, src, l), L2 = __shfl_sync(0xffffffffu, L, l), o2 = __shfl_sync(0xffffffffu, o, l);
        for (uint32_t j = lane; j < (L2 + 1) >> 1; j += 32) {
          const uint32_t k = s2 + 2 * j;
          const bool pair = 2 * j + 1 < L2;
          desc[o2 + j] = entries ? make_uint2(entries[k], pair ? entries[k + 1] : BA_NONE) : make_uint2(k, pair ? k + 1 : BA_NONE);
        }
      }
    }
  }

  template <class F>
  struct BaArgs {
    int round0;               // 1: operands are gathered from the base tables (descriptor = entry words)
    BasesSel<F> tables;       // round 0: base points (Montgomery form), one per selection (blockIdx.y)
    const Affine<F>* cur;     // later rounds: the previous round's output
    size_t cur_stride;        // elements per selection in cur
    Affine<F>* nxt;           // output
    size_t nxt_stride;
    const uint2* desc;        // per output slot
    const uint32_t* nslots;   // device: number of output slots (off_next[nb])
    F* prefix;                // per output slot: product of the thread's denominators in front of it (stride nxt_stride)
    int k;                    // slots per thread in this round (one inversion per thread and tile)
  };

  template <class T>
  __device__ __forceinline__ T ba_ldg(const T* p) // read-only 128-bit loads (inputs of a round are never written by it)
  {
    constexpr int NQ = sizeof(T) / 16;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    T r;
    uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      w[i] = __ldg(q + i);
    return r;
  }

  // tangent slope numerator 3x^2, out of line: doublings inside a bucket are rare (repeated base points)
  template <class F>
  __device__ __noinline__ F ba_tangent_num(const F& x)
  {
    F xx = x.sqr();
    return xx.dbl() + xx;
  }

  template <class F>
  struct BaView { // one selection's operands
    const Affine<F>* src; // base table (round 0) or previous output
    bool round0;
    __device__ __forceinline__ F x_of(uint32_t k) const { return ba_ldg(&src[round0 ? (k & 0x7fffffffu) : k].x); }
    __device__ __forceinline__ Affine<F> point(uint32_t k) const
    {
      Affine<F> p = ba_ldg(src + (round0 ? (k & 0x7fffffffu) : k));
      if (round0 && (k >> 31)) p.y = p.y.neg(); // -(0,0) stays the identity
      return p;
    }
  };

  // the denominator of a slot whose x coordinates coincide or vanish (doubling, cancellation, identity operand)
  template <class F>
  __device__ __noinline__ F ba_cold_den(const BaView<F>& V, uint2 d)
  {
    F den;
    pair_prepare(V.point(d.x), V.point(d.y), den);
    return den;
  }

  template <class F>
  __global__ void __launch_bounds__(BA_BLOCK, BaCfg<F>::kMinBlocks) ba_round_kernel(BaArgs<F> A)
  {
    const int K = A.k;
    const uint32_t S = *A.nslots;
    BaView<F> V;
    V.round0 = A.round0 != 0;
    V.src = V.round0 ? A.tables.p[blockIdx.y] : A.cur + (size_t)blockIdx.y * A.cur_stride;
    Affine<F>* nxt = A.nxt + (size_t)blockIdx.y * A.nxt_stride;
    F* prefix = A.prefix + (size_t)blockIdx.y * A.nxt_stride;
    const uint32_t tile_slots = (uint32_t)BA_BLOCK * (uint32_t)K;
    const uint32_t tiles = (S + tile_slots - 1) / tile_slots;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint32_t base = tile * tile_slots + threadIdx.x;
      // ---- forward: running product of the denominators (only the operands' x are read)
      F run = F::one();
#pragma unroll 1
      for (int i = 0; i < K; ++i) {
        const uint32_t slot = base + (uint32_t)i * BA_BLOCK;
        if (slot >= S) break;
        const uint2 d = __ldg(A.desc + slot);
        F den = F::one();
        if (d.y != BA_NONE) {
          const F ax = V.x_of(d.x), bx = V.x_of(d.y);
          den = bx - ax;
          if (den.is_zero() || ax.is_zero() || bx.is_zero()) den = ba_cold_den(V, d);
        }
        st_struct(prefix + slot, run);
        run = run * den;
      }
      F inv = batch_inverse(run);
      // ---- backward: recover each denominator's inverse, finish the addition
#pragma unroll 1
      for (int i = K - 1; i >= 0; --i) {
        const uint32_t slot = base + (uint32_t)i * BA_BLOCK;
        if (slot >= S) continue; // den was one: inv unchanged
        const uint2 d = __ldg(A.desc + slot);
        const Affine<F> a = V.point(d.x);
        if (d.y == BA_NONE) {
          st_struct(nxt + slot, a); // odd tail of its bucket: copied
          continue;
        }
        const Affine<F> b = V.point(d.y);
        F den;
        const int kind = pair_prepare(a, b, den);
        const F den_inv = ld_struct(prefix + slot) * inv;
        inv = inv * den;
        Affine<F> r;
        if (kind == PAIR_CHORD || kind == PAIR_TANGENT) {
          F num = b.y - a.y;
          if (kind == PAIR_TANGENT) num = ba_tangent_num(a.x);
          const F lam = num * den_inv;
          r.x = lam.sqr() - a.x - b.x;
          r.y = lam * (a.x - r.x) - a.y;
        } else {
          r = kind == PAIR_LEFT ? a : (kind == PAIR_RIGHT ? b : Affine<F>::inf());
        }
        st_struct(nxt + slot, r);
      }
    }
  }

  // after the last round: bucket b = sum of what is left in its segment (normally 1-3 points); segments longer than
  // BA_FIN_SERIAL go to the long list
  template <class F>
  __global__ void __launch_bounds__(128) ba_buckets_kernel(
    const uint32_t* off, const Affine<F>* cur, size_t pts_stride, int nb, XYZZ<F>* buckets, uint32_t* long_list, uint32_t* long_count)
  {
    cur += (size_t)blockIdx.y * pts_stride;
    buckets += (size_t)blockIdx.y * nb;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
      const uint32_t lo = off[b], hi = off[b + 1];
      if (lo == hi) continue; // empty bucket: never read by the reduction
      if (hi - lo > BA_FIN_SERIAL) {
        if (blockIdx.y == 0) long_list[atomicAdd(long_count, 1u)] = (uint32_t)b;
        continue;
      }
      XYZZ<F> acc = XYZZ<F>::from_affine(ld_struct(cur + lo));
      for (uint32_t k = lo + 1; k < hi; ++k)
        acc.madd(ld_struct(cur + k));
      st_struct(buckets + b, acc);
    }
  }

  template <class F>
  __global__ void __launch_bounds__(FOLD_BLOCK) ba_long_kernel(
    const uint32_t* off, const Affine<F>* cur, size_t pts_stride, int nb, XYZZ<F>* buckets, const uint32_t* long_list,
    const uint32_t* long_count)
  {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    cur += (size_t)blockIdx.y * pts_stride;
    buckets += (size_t)blockIdx.y * nb;
    const uint32_t nl = *long_count;
    for (uint32_t m = blockIdx.x; m < nl; m += gridDim.x) {
      const uint32_t b = long_list[m], lo = off[b], hi = off[b + 1];
      XYZZ<F> acc = XYZZ<F>::inf();
      for (uint32_t k = lo + threadIdx.x; k < hi; k += FOLD_BLOCK)
        acc.madd(ld_struct(cur + k));
      sh[threadIdx.x] = acc;
      __syncthreads();
      for (int s = FOLD_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
          XYZZ<F> a = sh[threadIdx.x];
          xyzz_add_ni(a, sh[threadIdx.x + s]);
          sh[threadIdx.x] = a;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) st_struct(buckets + b, sh[0]);
      __syncthreads();
    }
  }

  // Rounds of batched-affine accumulation for this plan.  B200_BATCH_AFFINE = 0: never (the XYZZ kernel); N: N rounds for
  // every MSM; unset / "auto": G2 only, floor(log2(average bucket length)) - 1 rounds when the buckets average >= 16
  // entries and the MSM is large enough to fill the GPU with tiles.  Measured on B200 at 3.2 M points x 13 windows
  // (profiles/r02_batch_affine.md): G2 20.9 ms against 22.6 ms (XYZZ); G1 9.4 ms against 8.7 ms - in G1 the round that
  // gathers from the base table is DRAM-bound (two gathers per add, ~180 B of DRAM traffic per 64 B point), so G1 stays
  // on the XYZZ kernel unless forced.
  static inline int msm_batch_affine_rounds(const MsmPlan& plan, bool g2)
  {
    static const int forced = [] {
      const char* e = getenv("B200_BATCH_AFFINE");
      if (!e || !*e || !strcmp(e, "auto")) return -1;
      int r = atoi(e);
      return r < 0 ? 0 : (r > 24 ? 24 : r);
    }();
    if (forced >= 0) return forced;
    if (!g2) return 0;
    const size_t E = plan.entries();
    if (E < ((size_t)1 << 21) || E >= ((size_t)1 << 31)) return 0;
    const size_t avg = E / (size_t)plan.nbuckets;
    if (avg < 16) return 0;
    int r = 0;
    while (((size_t)2 << r) <= avg) ++r;
    return r - 1;
  }

  // buckets[which * nb + b] = sum of the bucket's points for every selection, for all non-empty buckets
  template <class F>
  eIcicleError msm_accumulate_batched_enqueue(
    const MsmPlan& plan, const MsmSorted& sorted, const BasesSel<F>& sel, int nsel, int rounds, XYZZ<F>* buckets, cudaStream_t st)
  {
    const size_t E = plan.entries();
    const int nb = plan.nbuckets;
    if (E >= (1ull << 31) || rounds < 1 || nsel < 1 || nsel > MSM_MAX_SEL) return ICICLE_INVALID_ARGUMENT;
    // a round's output has at most ceil(E_in / 2) + nb slots (one copied tail per bucket)
    const size_t slots0 = (E + 1) / 2 + (size_t)nb, slots1 = (slots0 + 1) / 2 + (size_t)nb;
    const int scan_tiles = nb / 4096 + 4;
    static const int kmax = [] {
      const char* e = getenv("B200_BA_K");
      int k = e ? atoi(e) : BaCfg<F>::K;
      return k < 2 ? 2 : (k > 256 ? 256 : k);
    }();

    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_pts0 = 0;
    size_t o_pts1 = o_pts0 + al((size_t)nsel * slots0 * sizeof(Affine<F>));
    size_t o_prefix = o_pts1 + al((size_t)nsel * slots1 * sizeof(Affine<F>));
    size_t o_desc = o_prefix + al((size_t)nsel * slots0 * sizeof(F));
    size_t o_off0 = o_desc + al(slots0 * sizeof(uint2));
    size_t o_off1 = o_off0 + al(((size_t)nb + 1) * 4);
    size_t o_len = o_off1 + al(((size_t)nb + 1) * 4);
    size_t o_tiles = o_len + al(((size_t)nb + 1) * 4);
    size_t o_long = o_tiles + al((size_t)scan_tiles * 4);
    size_t total = o_long + al(((size_t)nb + 1) * 4);
    uint8_t* base = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&base, total, st), ICICLE_ALLOCATION_FAILED);
    Affine<F>* pts[2] = {(Affine<F>*)(base + o_pts0), (Affine<F>*)(base + o_pts1)};
    const size_t strides[2] = {slots0, slots1};
    uint32_t* offs[2] = {(uint32_t*)(base + o_off0), (uint32_t*)(base + o_off1)};
    uint2* desc = (uint2*)(base + o_desc);
    uint32_t* lens = (uint32_t*)(base + o_len);
    uint32_t* tiles = (uint32_t*)(base + o_tiles);
    uint32_t* long_list = (uint32_t*)(base + o_long); // [0] = count, list behind it
    cudaMemsetAsync(long_list, 0, 4, st);

    BaArgs<F> A;
    A.tables = sel;
    A.desc = desc;
    A.prefix = (F*)(base + o_prefix);
    const uint32_t* off = sorted.offsets;
    const Affine<F>* cur = pts[1];
    size_t cur_stride = slots1, in_bound = E;
    const int sms = sm_count();
    const size_t resident = (size_t)sms * BaCfg<F>::kMinBlocks;
    for (int r = 0; r < rounds; ++r) {
      uint32_t* off_next = offs[r & 1];
      B200_LAUNCH(ba_halve_kernel, grid_for(nb, 256, 8), 256, 0, st, off, nb, lens);
      msm_exclusive_scan(lens, nb, off_next, tiles, st);
      B200_LAUNCH(ba_desc_kernel, grid_for(nb, 256, 8), 256, 0, st, off, off_next, nb, r == 0 ? sorted.entries : nullptr, desc);
      const size_t out_bound = (in_bound + 1) / 2 + (size_t)nb;
      // slots per thread: whole waves of resident CTAs, at most kmax (one inversion per thread and tile)
      const size_t work = out_bound * (size_t)nsel, per_wave = resident * BA_BLOCK;
      const size_t waves = (work + per_wave * kmax - 1) / (per_wave * kmax);
      size_t k = (work + per_wave * waves - 1) / (per_wave * waves);
      if (k < 4) k = 4;
      const size_t ntiles = (out_bound + (size_t)BA_BLOCK * k - 1) / ((size_t)BA_BLOCK * k);
      A.round0 = r == 0;
      A.cur = cur;
      A.cur_stride = cur_stride;
      A.nxt = pts[r & 1];
      A.nxt_stride = strides[r & 1];
      A.nslots = off_next + nb;
      A.k = (int)k;
      B200_LAUNCH(ba_round_kernel<F>, dim3((unsigned)ntiles, (unsigned)nsel), BA_BLOCK, 0, st, A);
      off = off_next;
      cur = A.nxt;
      cur_stride = A.nxt_stride;
      if (out_bound < in_bound) in_bound = out_bound;
    }
    B200_LAUNCH(
      ba_buckets_kernel<F>, dim3(grid_for(nb, 128, 16), nsel), 128, 0, st, off, cur, cur_stride, nb, buckets, long_list + 1, long_list);
    B200_LAUNCH(
      ba_long_kernel<F>, dim3(sms, nsel), FOLD_BLOCK, FOLD_BLOCK * sizeof(XYZZ<F>), st, off, cur, cur_stride, nb, buckets,
      long_list + 1, long_list);
    cudaError_t ce = cudaGetLastError();
    cudaFreeAsync(base, st);
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] msm_accumulate_batched_enqueue: %s\n", cudaGetErrorString(ce));
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    return ICICLE_SUCCESS;
  }

} // namespace b200
