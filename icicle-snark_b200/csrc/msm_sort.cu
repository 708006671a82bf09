// Sort phase of the Pippenger MSM, independent of the group (G1/G2): signed-digit decomposition, bucket histogram,
// counting-sort scatter, work-item construction.  One sort can feed several accumulate/reduce phases that share
// the scalars (the prover's A, B1, C and B2 MSMs all use the witness; src/proof_helper.rs:198-206).
// See msm.cuh for the design table and the reference locations this replaces.
#include "msm.cuh"

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  static constexpr uint32_t DIGIT_NONE = 0xffffffffu;

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

  void msm_exclusive_scan(const uint32_t* in, int n, uint32_t* out, uint32_t* tile_sums, cudaStream_t st)
  {
    exclusive_scan(in, n, out, tile_sums, st);
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
        entries[pos] = (uint32_t)(i * pl.stride + w / pl.sets) | (d & 0x80000000u);
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // (4) work items: bucket k with n_k entries becomes ceil(n_k/T) items of <= T entries.
  static __global__ void __launch_bounds__(256) msm_item_count_kernel(MsmDev pl, const uint32_t* offsets, uint32_t* nitems)
  {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < pl.nbuckets; k += gridDim.x * blockDim.x) {
      uint32_t cnt = offsets[k + 1] - offsets[k];
      nitems[k] = (cnt + pl.item_cap - 1) / pl.item_cap;
    }
  }

  // writes the items of every bucket (unsorted, bucket-major) + the histogram of item lengths;
  // buckets with >1 item are appended to the `multi` list.
  static __global__ void __launch_bounds__(256) msm_item_build_kernel(
    MsmDev pl, const uint32_t* offsets, const uint32_t* item_off, MsmItem* items, uint32_t* len_hist, uint32_t* multi,
    uint32_t* multi_count)
  {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < pl.nbuckets; k += gridDim.x * blockDim.x) {
      uint32_t beg = offsets[k], cnt = offsets[k + 1] - beg;
      uint32_t first = item_off[k], ni = item_off[k + 1] - first;
      if (ni == 0) continue; // empty bucket: the reduction phase reads it as the identity
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

  MsmDev msm_dev_plan(const MsmPlan& plan)
  {
    MsmDev pl;
    pl.n = plan.n;
    pl.c = plan.c;
    pl.windows = plan.windows;
    pl.factor = plan.factor;
    pl.stride = plan.stride;
    pl.sets = plan.sets;
    pl.bpw = plan.bpw;
    pl.nbuckets = plan.nbuckets;
    pl.item_cap = plan.item_cap;
    for (int i = 0; i < 9; ++i)
      pl.h[i] = plan.hconst[i];
    return pl;
  }

  eIcicleError msm_sort_enqueue(const MsmPlan& plan, const Fr* scalars, bool scalars_mont, MsmSorted* out, cudaStream_t st)
  {
    MsmDev pl = msm_dev_plan(plan);
    const size_t ne = plan.entries();
    const int nb = plan.nbuckets;
    const size_t max_items = (size_t)nb + ne / plan.item_cap + 1;
    const int scan_tiles_b = (nb + SCAN_TILE - 1) / SCAN_TILE + 2;

    // one stream-ordered scratch block, carved up (the pool keeps it across calls)
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
    size_t o_small = o_lenoff + al((size_t)(plan.item_cap + 2) * 4); // [0] = multi_count
    size_t o_multi = o_small + 256;
    size_t o_items = o_multi + al((size_t)nb * 4);
    size_t o_sorted = o_items + al(max_items * sizeof(MsmItem));
    size_t total = o_sorted + al(max_items * sizeof(MsmItem));

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

    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t e) {
      if (ce == cudaSuccess) ce = e;
    };
    chk(cudaMemsetAsync(hist, 0, o_offsets - o_hist, st));
    chk(cudaMemsetAsync(len_hist, 0, o_multi - o_lenhist, st)); // len_hist, len_off, multi_count

    B200_LAUNCH(msm_digits_kernel, grid_for(plan.n, 256, 8), 256, 0, st, pl, scalars, scalars_mont, digits, hist);
    exclusive_scan(hist, nb, offsets, tiles, st);
    chk(cudaMemcpyAsync(cursor, offsets, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st));
    B200_LAUNCH(msm_scatter_kernel, grid_for(ne, 256, 8), 256, 0, st, pl, digits, cursor, entries);
    B200_LAUNCH(msm_item_count_kernel, grid_for(nb, 256, 8), 256, 0, st, pl, offsets, nitems);
    exclusive_scan(nitems, nb, item_off, tiles, st); // item_off[nb] = total items (device-side only)
    B200_LAUNCH(msm_item_build_kernel, grid_for(nb, 256, 8), 256, 0, st, pl, offsets, item_off, items, len_hist, multi, multi_count);
    exclusive_scan(len_hist, plan.item_cap + 1, len_off, tiles, st);
    B200_LAUNCH(msm_item_sort_kernel, grid_for(max_items, 256, 8), 256, 0, st, items, item_off + nb, plan.item_cap, len_off, sorted);
    chk(cudaGetLastError());
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] msm_sort_enqueue: %s\n", cudaGetErrorString(ce));
      cudaFreeAsync(base, st);
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    out->base = base;
    out->entries = entries;
    out->offsets = offsets;
    out->item_off = item_off;
    out->multi = multi;
    out->multi_count = multi_count;
    out->sorted = sorted;
    out->max_items = max_items;
    return ICICLE_SUCCESS;
  }

  void msm_sorted_free(MsmSorted* s, cudaStream_t st)
  {
    if (s && s->base) cudaFreeAsync(s->base, st);
    if (s) s->base = nullptr;
  }

} // namespace b200
