// EXPERIMENTAL, OFF BY DEFAULT (B200_BATCH_AFFINE=<rounds> turns it on): bucket accumulation by batched affine
// addition - kernels around the per-thread bodies of batch_affine.cuh.  The bodies are validated on the host
// (batch_affine_model.cu, tests/test_host_math.py); these wrappers and the driver below have been compiled for
// sm_100a but NOT yet run on hardware (the round's GPU budget was spent) - no measurement or parity claim rests on
// them, and with the variable unset msm_reduce_enqueue never reaches this file's code.  DESIGN.md section 8.1.
//
// Replaces msm_accumulate_kernel + the fold kernels for one group of MSMs sharing a sort: `rounds` pairwise-tree
// rounds (three passes each), then one thread per bucket converts what is left of its segment to XYZZ.
#pragma once
#include "batch_affine.cuh"

namespace b200 {

  // len_next[b] = ceil(len[b] / 2)
  static __global__ void __launch_bounds__(256) ba_halve_kernel(const uint32_t* off, int nb, uint32_t* len_next)
  {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
      uint32_t L = off[b + 1] - off[b];
      len_next[b] = (L + 1) >> 1;
    }
  }

  template <class F>
  __global__ void __launch_bounds__(128) ba_prefix_kernel(BaLaunch<F> L)
  {
    BaRound<F> R = ba_round_of(L, blockIdx.y);
    const uint32_t threads = (R.off_next[R.nb] + BA_M - 1) / BA_M;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < threads; t += gridDim.x * blockDim.x)
      ba_prefix_thread(R, t);
  }

  template <class F>
  __global__ void __launch_bounds__(128) ba_invert_kernel(BaLaunch<F> L)
  {
    BaRound<F> R = ba_round_of(L, blockIdx.y);
    const uint32_t n_totals = (R.off_next[R.nb] + BA_M - 1) / BA_M;
    const uint32_t threads = (n_totals + BA_M2 - 1) / BA_M2;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < threads; u += gridDim.x * blockDim.x)
      ba_invert_thread(R.totals, n_totals, u);
  }

  template <class F>
  __global__ void __launch_bounds__(128) ba_finish_kernel(BaLaunch<F> L)
  {
    BaRound<F> R = ba_round_of(L, blockIdx.y);
    const uint32_t threads = (R.off_next[R.nb] + BA_M - 1) / BA_M;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < threads; t += gridDim.x * blockDim.x)
      ba_finish_thread(R, t);
  }

  template <class F>
  __global__ void __launch_bounds__(128)
    ba_buckets_kernel(const uint32_t* off, const Affine<F>* cur, size_t pts_stride, int nb, XYZZ<F>* buckets)
  {
    cur += (size_t)blockIdx.y * pts_stride;
    buckets += (size_t)blockIdx.y * nb;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
      XYZZ<F> acc;
      if (ba_bucket_thread(off, cur, b, acc)) st_struct(buckets + b, acc);
    }
  }

  // rounds requested through B200_BATCH_AFFINE (0 = the XYZZ accumulate path; read once)
  static inline int batch_affine_rounds()
  {
    static const int rounds = [] {
      const char* e = getenv("B200_BATCH_AFFINE");
      int r = e ? atoi(e) : 0;
      return r < 0 ? 0 : (r > 32 ? 32 : r);
    }();
    return rounds;
  }

  // launches one pass per call; grids are sized with the host-side slot bounds, the kernels read the true counts
  template <class F>
  struct BaDeviceExec {
    cudaStream_t st;
    uint32_t* lens;
    uint32_t* tiles;
    void next_offsets(const uint32_t* off, int nb, uint32_t* off_next)
    {
      B200_LAUNCH(ba_halve_kernel, grid_for(nb, 256, 8), 256, 0, st, off, nb, lens);
      msm_exclusive_scan(lens, nb, off_next, tiles, st);
    }
    void prefix(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      B200_LAUNCH(ba_prefix_kernel<F>, dim3(grid_for(threads, 128, 16), nsel), 128, 0, st, L);
    }
    void invert(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      B200_LAUNCH(ba_invert_kernel<F>, dim3(grid_for(threads, 128, 16), nsel), 128, 0, st, L);
    }
    void finish(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      B200_LAUNCH(ba_finish_kernel<F>, dim3(grid_for(threads, 128, 16), nsel), 128, 0, st, L);
    }
  };

  // buckets[which * nb + b] = sum of the bucket's points for every selection, for all non-empty buckets
  template <class F>
  eIcicleError msm_accumulate_batched_enqueue(
    const MsmPlan& plan, const MsmSorted& sorted, const BasesSel<F>& sel, int nsel, int rounds, XYZZ<F>* buckets, cudaStream_t st)
  {
    static_assert(BA_MAX_SEL == MSM_MAX_SEL, "selection planes");
    const size_t E = plan.entries();
    const int nb = plan.nbuckets;
    if (E >= (1ull << 31) || rounds < 1 || nsel < 1 || nsel > BA_MAX_SEL) return ICICLE_INVALID_ARGUMENT;
    const size_t slots0 = ba_slot_bound(E, nb), thr0 = ba_threads_for(slots0);
    const int scan_tiles = nb / 4096 + 4;

    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_pts0 = 0;
    size_t o_pts1 = o_pts0 + al((size_t)nsel * slots0 * sizeof(Affine<F>));
    size_t o_prefix = o_pts1 + al((size_t)nsel * slots0 * sizeof(Affine<F>));
    size_t o_totals = o_prefix + al((size_t)nsel * slots0 * sizeof(F));
    size_t o_off0 = o_totals + al((size_t)nsel * thr0 * sizeof(F));
    size_t o_off1 = o_off0 + al(((size_t)nb + 1) * 4);
    size_t o_len = o_off1 + al(((size_t)nb + 1) * 4);
    size_t o_tiles = o_len + al(((size_t)nb + 1) * 4);
    size_t total = o_tiles + al((size_t)scan_tiles * 4);
    uint8_t* base = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&base, total, st), ICICLE_ALLOCATION_FAILED);

    BaDeviceExec<F> ex{st, (uint32_t*)(base + o_len), (uint32_t*)(base + o_tiles)};
    BaResult res = ba_run_rounds<F>(
      ex, E, nb, nsel, rounds, sorted.entries, sel.p, sorted.offsets, (Affine<F>*)(base + o_pts0), (Affine<F>*)(base + o_pts1),
      (F*)(base + o_prefix), (F*)(base + o_totals), (uint32_t*)(base + o_off0), (uint32_t*)(base + o_off1));
    B200_LAUNCH(
      ba_buckets_kernel<F>, dim3(grid_for(nb, 128, 16), nsel), 128, 0, st, res.off, (const Affine<F>*)res.cur, res.pts_stride, nb,
      buckets);
    cudaError_t ce = cudaGetLastError();
    cudaFreeAsync(base, st);
    if (ce != cudaSuccess) {
      fprintf(stderr, "[icicle_b200] msm_accumulate_batched_enqueue: %s\n", cudaGetErrorString(ce));
      return translate(ce, ICICLE_UNKNOWN_FALLBACK);
    }
    return ICICLE_SUCCESS;
  }

} // namespace b200
