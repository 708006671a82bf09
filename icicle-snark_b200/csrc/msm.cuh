// Pippenger MSM for BN254 G1 (F = Fq) and G2 (F = Fq2), sm_100a.
//
// Replaces /root/reference/icicle/backend/cuda/src/msm/cuda_msm.cuh:960-1127 (bucket_method_msm) and
// its kernels (:156-387).  Same mathematical contract (sum_i s_i * P_i), different algorithm:
//
//   reference                                   here
//   ---------                                   ----
//   unsigned c-bit digits (:166-203)            signed digits (add-constant recoding): half the buckets
//   CUB radix sort of all 32 key bits + RLE     counting sort: histogram in the digit pass, one scan,
//   + scan + second sort by bucket size         window-major scatter (L2-resident write window)
//   one thread per bucket, serial (:223-255)    one thread per fixed-length SEGMENT of the sorted entry
//   + large-bucket side path (:666-765)         list: perfectly balanced for any scalar distribution;
//                                               buckets cut by a segment border are stitched by a
//                                               fix-up pass (short spans: 1 thread, long spans: 1 CTA)
//   projective RCB adds (12M+)                  XYZZ buckets, affine bases: 8M+2S per accumulate step
//   log-halving reduction (:846-942)            chunked running sums + per-set tree, Horner over sets
//   precompute_factor (:29-43)                  same idea; factor == #windows collapses all windows
//                                               into ONE bucket set (no doublings at all) — affordable
//                                               because the zkey bases live in 180 GB of HBM
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "staging.cuh"

namespace b200 {

  struct MsmPlan {
    int n;           // points
    int c;           // window bits
    int windows;     // W = ceil((bitsize+1)/c)
    int factor;      // precompute factor f (tables of 2^(c*sets*j) * P)
    int sets;        // bucket sets = ceil(W/f)
    int buckets;     // per set: 2^(c-1)
    int keys;        // sets * buckets
    int seg;         // entries per accumulate thread
    uint32_t hconst[9]; // recoding constant H = sum_w (2^(c-1)-1) 2^(cw), 288 bits
    size_t max_entries() const { return (size_t)n * windows; }
    size_t max_segments() const { return (max_entries() + seg - 1) / seg; }
  };

  MsmPlan make_msm_plan(int n, int c_req, int bitsize, int factor, bool g2);

  // Enqueue one MSM on `st`.  All pointers are DEVICE pointers; `bases_mont` holds factor*n affine
  // points in Montgomery form laid out table-major ([j][i]); `out_xyzz` receives the result (Montgomery).
  template <class F>
  eIcicleError msm_enqueue(
    const MsmPlan& plan, const Fr* scalars, bool scalars_mont, const Affine<F>* bases_mont, XYZZ<F>* out_xyzz,
    cudaStream_t st);

  // out[j][i] = 2^(shift*j) * in[i], affine Montgomery in and out
  template <class F>
  eIcicleError precompute_enqueue(const Affine<F>* in, int n, int factor, int shift, Affine<F>* out, cudaStream_t st);

  // XYZZ (Montgomery) -> reference projective layout, standard form
  template <class F>
  eIcicleError xyzz_to_projective_enqueue(const XYZZ<F>* in, int count, Projective<F>* out_std, cudaStream_t st);

  // number of kernel launches the last msm_enqueue issued (bench.py's gpu_launches claim)
  int msm_last_launch_count();

} // namespace b200
