// Pippenger MSM for BN254 G1 (F = Fq) and G2 (F = Fq2), sm_100a.
//
// Replaces /root/reference/icicle/backend/cuda/src/msm/cuda_msm.cuh:960-1127 (bucket_method_msm) and
// its kernels (:156-387).  Same mathematical contract (sum_i s_i * P_i), different algorithm:
//
//   reference                                   here
//   ---------                                   ----
//   unsigned c-bit digits (:166-203)            signed digits by add-constant recoding: half the buckets,
//                                               every window independent (no carry between windows)
//   CUB radix sort of all 32 key bits + RLE     counting sort: histogram in the digit pass, one scan,
//   + scan + second sort by bucket size         window-major scatter (L2-resident write window); bucket
//                                               work items counting-sorted by length so a warp's lanes finish together
//   one thread per bucket, serial (:223-255)    one thread per work item = (bucket, <= T entries); buckets longer
//   + large-bucket side path (:666-765)         than T (skewed scalars) are cut into items whose partial sums
//                                               are folded by a second small kernel
//   projective RCB adds (12M+)                  XYZZ buckets, affine bases: 8M+2S per accumulate step
//   log-halving reduction (:846-942)            chunked running sums, one CTA tree per window, Horner over windows
//   precompute_factor (:29-43)                  same table layout out[i*f+j] = 2^(shift*j) P_i; window width chosen so the
//                                               windows fit the factor (one bucket set, no Horner tail)
//   one sort per MSM                            the sort (msm_sort.cu) is group-independent and can feed several
//                                               accumulate/reduce phases sharing the scalars: the prover's A, B1, C
//                                               run as ONE three-table G1 launch, B2 reuses the same sort
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "staging.cuh"

namespace b200 {

  struct MsmPlan {
    int n;        // scalars
    int c;        // window bits
    int windows;  // digits per scalar: ceil((bitsize+2)/c)
    int factor;   // precomputed multiples used per point: min(f, windows)
    int stride;   // table stride: entry (i, j) of the base table is bases[i * stride + j] (the caller's f)
    int sets;     // bucket sets after precompute folding: ceil(windows/f)
    int bpw;      // buckets per set: 2^(c-1)
    int nbuckets; // sets * bpw
    int item_cap; // T: max entries per accumulate work item
    uint32_t hconst[9]; // recoding constant H = sum_w 2^(c-1) 2^(cw), 288 bits
    size_t entries() const { return (size_t)n * windows; }
  };

  MsmPlan make_msm_plan(int n, int c_req, int bitsize, int factor, bool g2);

  struct MsmDev { // plan fields the kernels need, passed by value
    int n, c, windows, factor, stride, sets, bpw, nbuckets, item_cap;
    uint32_t h[9];
  };
  MsmDev msm_dev_plan(const MsmPlan& plan);

  struct MsmItem {
    uint32_t begin;  // first entry
    uint32_t len;    // 1..T
    uint32_t bucket; // bucket key
    uint32_t dst;    // index into `buckets` (single-item bucket) or 0x80000000|index into `partials`
  };

  // Result of the group-independent sort phase (msm_sort.cu): bucket-sorted point references + work items.
  struct MsmSorted {
    uint8_t* base = nullptr;       // scratch block (stream-ordered allocation), freed by msm_sorted_free
    uint32_t* entries = nullptr;   // (point index | sign << 31), bucket-sorted
    uint32_t* offsets = nullptr;   // nbuckets + 1 exclusive offsets into entries
    uint32_t* item_off = nullptr;  // nbuckets + 1; item_off[nbuckets] = number of work items (device side)
    uint32_t* multi = nullptr;     // keys of buckets cut into several items
    uint32_t* multi_count = nullptr;
    MsmItem* sorted = nullptr;     // work items, longest first
    size_t max_items = 0;
  };
  // out[0..n] = exclusive scan of in[0..n), out[n] = total; tile_sums: >= n / 4096 + 2 words of scratch (msm_sort.cu)
  void msm_exclusive_scan(const uint32_t* in, int n, uint32_t* out, uint32_t* tile_sums, cudaStream_t st);
  eIcicleError msm_sort_enqueue(const MsmPlan& plan, const Fr* scalars, bool scalars_mont, MsmSorted* out, cudaStream_t st);
  void msm_sorted_free(MsmSorted* s, cudaStream_t st);

  // Accumulate + reduce `nsel` (<= 4) MSMs that share `sorted` (same scalars, same plan) over different base-point
  // arrays; out_std[k] receives MSM k in the reference's boundary layout.
  // `gate` (optional): the bucket-accumulation phase waits for this event (the scratch allocation before it does not)
  template <class F>
  eIcicleError msm_reduce_enqueue(
    const MsmPlan& plan, const MsmSorted& sorted, const Affine<F>* const* bases_mont, int nsel, Projective<F>* out_std,
    cudaStream_t st, cudaEvent_t gate = nullptr);

  // Enqueue one MSM on `st`.  All pointers are DEVICE pointers; `bases_mont` holds n*factor affine
  // points in Montgomery form ([i*f + j] = 2^(shift*j) P_i); `out_std` receives the result in the
  // reference's boundary layout (homogeneous projective, standard form).
  template <class F>
  eIcicleError msm_enqueue(
    const MsmPlan& plan, const Fr* scalars, bool scalars_mont, const Affine<F>* bases_mont, Projective<F>* out_std,
    cudaStream_t st);

  // out[i*f + j] = 2^(shift*j) * in[i]; in: affine (standard or Montgomery), out: affine Montgomery or standard
  template <class F>
  eIcicleError precompute_enqueue(
    const Affine<F>* in, bool in_mont, int n, int factor, int shift, Affine<F>* out, bool out_mont, cudaStream_t st);

  // kernel launches issued by this library since load (bench.py's gpu_launches claim)
  extern unsigned long long g_launches;
  // Profiling mode (b200_profile_accumulate / b200_profile_records; bench.py's roofline): while enabled, the bucket-
  // accumulation phase of every MSM is isolated by device-wide synchronisation on both sides and timed with CUDA events
  // on its own stream, so each record is the duration of that launch running ALONE (the proof's streams otherwise
  // overlap it with other kernels).  Off by default; never enabled inside a timed region.
  struct MsmProfileRec {
    int g2, nsel, n, windows, c, factor, nbuckets, batched; // batched: rounds of batched-affine accumulation (0 = XYZZ)
    float ms;
  };
  extern int g_profile_mode;
  void msm_profile_begin(cudaStream_t st);
  void msm_profile_end(cudaStream_t st, const MsmPlan& plan, int g2, int nsel, int batched);

} // namespace b200
