// Fused Groth16 proving path: ZKeyCache (device-resident, Montgomery-form zkey), R1CS evaluation, the
// quotient chain, the five MSMs, the blinding epilogue and the proof.json/public.json writers.
//
// Replaces, behind the b200_groth16_* entry points of include/icicle_b200.h:
//   /root/reference/src/cache.rs:58-72,117-256,264-289      ZKeyCache / CacheManager
//   /root/reference/src/proof_helper.rs:31-170              construct_r1cs
//   /root/reference/src/proof_helper.rs:172-241             groth16_commitments
//   /root/reference/src/proof_helper.rs:243-317             groth16_prove_helper (epilogue)
//   /root/reference/src/file_wrapper.rs:45-103,169-237, src/zkey.rs:47-85, src/conversions.rs:30-56, src/lib.rs:33-61
//
// What changes relative to the reference's call sequence (same results):
//   * zkey points stay in the Montgomery form the file already stores them in (the reference converts
//     them out, cache.rs:208-212); coefficients stay as stored (coef*R^2): a Montgomery product with a
//     standard-form witness value is then directly the Montgomery form of coef*w.
//   * A*w, B*w are evaluated on the GPU from a CSR built once per zkey (the reference gathers on the host,
//     crosses PCIe twice and scatters in a single-thread loop, proof_helper.rs:55-99).
//   * iNTT -> x keys -> NTT runs on the Stockham passes of ntt.cu with 1/N * keys fused into the last
//     iNTT pass; A'.B' - C' and the conversion to standard form are one kernel.
//   * A, B1, C, B2 take the same scalars: one digit decomposition + counting sort feeds ONE three-table G1 launch
//     (s_g1) and the G2 launch (s_g2); they start as soon as the witness is on the device, concurrently with the
//     quotient chain, and H follows on the chain's (higher-priority) stream.  B columns at infinity are dropped at
//     build time (sparse-B path).  No host synchronisation until the five results are back; the blinding multiples
//     of delta are computed on the host meanwhile.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <vector>

#include "host_math.h"
#include "comm.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "staging.cuh"

namespace b200 {

#define B200_LAUNCH(kernel, grid, block, smem, st, ...)                                                                \
  do {                                                                                                                 \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                                            \
    ++g_launches;                                                                                                      \
  } while (0)

  // ------------------------------------------------------------------------------------------------ binfile
  struct Section {
    const uint8_t* p = nullptr;
    uint64_t size = 0;
  };

  // iden3 binfile container (file_wrapper.rs:45-103): magic, u32 version, u32 n_sections, then
  // {u32 id, u64 len, payload}*.  Returns false on malformed input.
  static bool parse_binfile(const uint8_t* buf, size_t len, const char* magic, uint32_t max_version, std::map<uint32_t, Section>& out)
  {
    if (len < 12 || memcmp(buf, magic, 4) != 0) return false;
    uint32_t version, nsec;
    memcpy(&version, buf + 4, 4);
    memcpy(&nsec, buf + 8, 4);
    if (version > max_version) return false;
    size_t pos = 12;
    for (uint32_t i = 0; i < nsec; ++i) {
      if (pos + 12 > len) return false;
      uint32_t id;
      uint64_t sz;
      memcpy(&id, buf + pos, 4);
      memcpy(&sz, buf + pos + 4, 8);
      pos += 12;
      if (sz > len - pos) return false;
      if (!out.count(id)) out[id] = {buf + pos, sz}; // first occurrence wins (sections[id][0])
      pos += sz;
    }
    return true;
  }

  static const uint32_t FR_MODULUS[8] = {0xf0000001, 0x43e1f593, 0x79b97091, 0x2833e848,
                                         0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};
  static const uint32_t FQ_MODULUS[8] = {0xd87cfd47, 0x3c208c16, 0x6871ca8d, 0x97816a91,
                                         0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};

  // ------------------------------------------------------------------------------------------------ kernels
  // rows: A row t and B row t evaluated by one thread; out layout as the reference's d_vec:
  // d[0..N) = B.w, d[N..2N) = A.w, d[2N..3N) = A.w * B.w   (proof_helper.rs:94-114), Montgomery form.
  // val = coef*R^2 (as stored in the zkey), w standard form: val (x) w = coef*w*R.
  static __global__ void __launch_bounds__(256) r1cs_eval_kernel(
    const uint32_t* row_ptr, const uint32_t* col, const Fr* val, const Fr* witness, uint32_t N, Fr* d)
  {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
      Fr acc[2];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        uint32_t beg = row_ptr[m * N + t], end = row_ptr[m * N + t + 1];
        Fr a = Fr::zero();
        for (uint32_t e = beg; e < end; ++e)
          a = a + ld_fr(val + e) * ld_fr(witness + col[e]);
        acc[m] = a;
      }
      st_fr(d + N + t, acc[0]);
      st_fr(d + t, acc[1]);
      st_fr(d + 2 * (size_t)N + t, acc[0] * acc[1]);
    }
  }

  // h[i] = from_mont(a[i]*b[i] - c[i])  (proof_helper.rs:153-167, plus the conversion the MSM digits need)
  static __global__ void __launch_bounds__(256)
    quotient_combine_kernel(const Fr* a, const Fr* b, const Fr* c, uint32_t count, Fr* h)
  {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      Fr v = ld_fp_coherent(a + i) * ld_fp_coherent(b + i) - ld_fp_coherent(c + i);
      st_fr(h + i, Fr::from_mont(v));
    }
  }

  // wb[k] = w[idx[k]]: witness values of the signals that have a B point
  static __global__ void __launch_bounds__(256) gather_scalars_kernel(const Fr* w, const uint32_t* idx, uint32_t n, Fr* out)
  {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
      st_fr(out + k, ld_fr(w + idx[k]));
  }

  struct PowTab {
    Fr pw[30];
  };
  // keys[i] = scale * g^i  (g^(2^j) table, Montgomery).  Replaces the host loop of cache.rs:264-289.
  static __global__ void __launch_bounds__(256) powers_kernel(PowTab t, Fr scale, int logn, Fr* out)
  {
    size_t n = (size_t)1 << logn;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      Fr acc = scale;
      for (int b = 0; b < logn; ++b)
        if ((i >> b) & 1) acc = acc * t.pw[b];
      st_fr(out + i, acc);
    }
  }

} // namespace b200

using namespace b200;

// ---------------------------------------------------------------------------------------------------- cache
struct b200_zkey_cache {
  int device = 0;
  int rank = 0, world = 1;
  uint32_t n_vars = 0, n_public = 0, domain_size = 0, power = 0;
  uint64_t n_coef = 0, device_bytes = 0;
  int precompute = 1;
  // verification-key points needed by the epilogue (Montgomery, as stored: zkey.rs:61-69)
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
  // this rank's [lo,hi) of every base-point section (b200_shard_plan): 0 = H, 1 = A, 2 = B1, 3 = C, 4 = B2
  uint32_t sec_lo[5] = {0, 0, 0, 0, 0}, sec_hi[5] = {0, 0, 0, 0, 0};
  uint32_t h_lo = 0, h_hi = 0; // == section 0
  int plan_mode = 0;           // shard_plan mode and skew this cache was cut with
  double plan_skew = 0;
  uint32_t w_lo = 0, w_hi = 0; // span of the witness this rank's MSMs read
  G1Affine* pH = nullptr;
  MsmPlan planH;
  // The witness MSMs (A, B1, C in G1, B2 in G2) take the same scalars over signal-indexed tables: sections that cover
  // the same signal range on this rank form a GROUP with one digit decomposition + counting sort feeding one fused G1
  // accumulate/reduce over its G1 tables and one G2 accumulate/reduce.  One GPU: a single group {A, B1, C, B2}.
  // B1/B2 columns that are points at infinity (signals absent from every B row: the norm in circom circuits) are
  // dropped at build time when they are >= 1/8 of the range: such a COMPACT group keeps `idx` (surviving signals
  // relative to lo) and gathers their witness values into d_w before its sort.
  struct WGroup {
    uint32_t lo = 0, hi = 0, n = 0; // signal range; points per table (hi - lo, or the kept count when compact)
    bool compact = false;
    uint32_t* idx = nullptr;
    Fr* d_w = nullptr;
    MsmPlan plan;
    int n_g1 = 0;
    G1Affine* g1[3] = {nullptr, nullptr, nullptr};
    int g1_slot[3] = {0, 0, 0}; // result slot of each table: 0 = A, 1 = B1, 2 = C
    G2Affine* g2 = nullptr;
    G1Projective* d_out = nullptr; // results of the fused G1 launch before they move to their slots
    cudaEvent_t ev_sort = nullptr, ev_done = nullptr;
  };
  std::vector<WGroup> groups;
  uint32_t n_b = 0, b_total = 0; // B1 points kept / B1 range size on this rank (b200_zkey_cache_b_points)
  // R1CS in CSR over rows [A rows 0..N) | B rows 0..N)]
  uint32_t *row_ptr = nullptr, *col = nullptr;
  Fr* val = nullptr;
  Fr* keys = nullptr; // 1/N * w_2N^i, Montgomery
  // per-proof workspace
  Fr *d_witness = nullptr, *d_vec = nullptr, *d_h = nullptr;
  uint8_t* d_results = nullptr; // 4 x G1 projective + 1 x G2 projective
  uint8_t* h_results = nullptr; // pinned
  cudaStream_t s_copy = nullptr, s_g1 = nullptr, s_g2 = nullptr, s_g3 = nullptr, s_q = nullptr, s_h = nullptr;
  cudaEvent_t ev_start = nullptr, ev_h2d = nullptr, ev_r1cs = nullptr, ev_ntt = nullptr, ev_g1 = nullptr,
              ev_g2 = nullptr, ev_q = nullptr, ev_prev = nullptr, ev_b1 = nullptr, ev_free = nullptr;
  std::mutex mu;
  bool in_flight = false; // commit_begin succeeded and holds `mu` until commit_end
  // in-library exchange of the sharded prover (b200_groth16_prove_sharded): this rank's slice of the three transformed
  // polynomials, every rank's partial sums
  Fr* qx_slices = nullptr;
  uint8_t *d_all_parts = nullptr, *h_all_parts = nullptr;
  cudaEvent_t ev_slice = nullptr, ev_xch = nullptr;
  uint8_t* h_stage = nullptr; // pinned staging for a pageable host witness (copy_pageable), n_vars * 32 B, lazily allocated
  size_t wit_slice = 0; // d_witness holds world * wit_slice elements (>= n_vars): in-place all-gather of the uploaded slices
};

namespace b200 {

  template <class T>
  static cudaError_t dev_alloc(T** p, size_t count, b200_zkey_cache* c)
  {
    size_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e == cudaSuccess) c->device_bytes += bytes;
    return e;
  }

  static void shard(uint32_t n, int rank, int world, uint32_t* lo, uint32_t* hi)
  {
    *lo = (uint32_t)((uint64_t)n * rank / world);
    *hi = (uint32_t)((uint64_t)n * (rank + 1) / world);
  }

  // quotient polynomials rank `rank` transforms when the chain is split (must match multi_gpu.poly_owner)
  static int polys_owned(int rank, int world)
  {
    if (world >= 3) return rank < 3 ? 1 : 0;
    if (world == 2) return rank == 0 ? 2 : 1;
    return 3;
  }

  // Witness-MSM shard of `rank` when the quotient chain is split across ranks: the polynomial owners carry the
  // transforms on top of their MSM shard, so they get a smaller one.  share(r) = 1/world + (3/world - polys(r)) * skew,
  // skew = (time of one polynomial's iNTT + NTT) / (time of all witness MSMs on one GPU); skew = 0 is the equal split.
  static void shard_skewed(uint32_t n, int rank, int world, double skew, uint32_t* lo, uint32_t* hi)
  {
    if (!(skew > 0) || world < 2) return shard(n, rank, world, lo, hi);
    if (skew > 0.2) skew = 0.2;
    double acc = 0, bounds[65];
    if (world > 64) return shard(n, rank, world, lo, hi);
    bounds[0] = 0;
    for (int r = 0; r < world; ++r) {
      double share = 1.0 / world + (3.0 / world - polys_owned(r, world)) * skew;
      if (share <= 0) return shard(n, rank, world, lo, hi); // too few ranks for this skew: equal split
      acc += share;
      bounds[r + 1] = acc;
    }
    auto at = [&](int r) {
      if (r <= 0) return (uint32_t)0;
      if (r >= world) return n;
      uint64_t v = (uint64_t)((double)n * (bounds[r] / acc) + 0.5);
      return (uint32_t)(v > n ? n : v);
    };
    *lo = at(rank);
    *hi = at(rank + 1);
    if (*hi < *lo) *hi = *lo;
  }

  // ---- which part of which base-point section a rank holds; sections: 0 = H, 1 = A, 2 = B1, 3 = C, 4 = B2
  // mode 0 ("uniform"): every section cut into `world` contiguous ranges - H equally, the four signal-indexed sections
  //   with the polynomial owners' skew (shard_skewed).  Every rank runs five small MSMs.
  // mode 1 ("line"): the five sections laid end to end, weighted by their cost per point (G1 = 1, G2 = w2), with the
  //   quotient-polynomial transforms charged to their owners (w_ntt * N point-equivalents per polynomial), cut into
  //   `world` pieces of equal cost (B200_PLAN_W2 / B200_PLAN_WH / B200_PLAN_WNTT override the fitted weights).  A rank holds one or two LARGE pieces (whole tables where possible): it keeps the
  //   wide windows of the single-GPU plan (13 digits per scalar instead of 15 at 3200k constraints on 8 GPUs) and runs two
  //   or three bucket reductions instead of five.  Cuts that fall within 4 % of a section edge snap to the edge.
  struct ShardPlan {
    uint32_t lo[5], hi[5];
  };
  // Default cut: B200_SHARD_PLAN = uniform | line when set; otherwise what measured faster at 3200k constraints on B200s
  // (ms per proof, line vs uniform: 2 GPUs 30.5 vs 31.0, 4 GPUs 18.4 vs 18.0, 8 GPUs 10.2 vs 10.5 - profiles/r02_multi_gpu.md):
  // few ranks hold whole tables under the line cut and keep the shared sort of equal ranges, many ranks gain from the
  // wide windows; in between the unfused pieces of a rank (one sort each) cost more than the windows save.
  static int plan_mode_default(int world)
  {
    if (world < 2) return 0;
    const char* e = getenv("B200_SHARD_PLAN");
    if (!e || !*e) return (world == 2 || world >= 6) ? 1 : 0;
    return (e[0] == '0' || e[0] == 'u' || e[0] == 'U') ? 0 : 1;
  }
  static double env_double(const char* name, double dflt)
  {
    const char* e = getenv(name);
    return (e && *e) ? atof(e) : dflt;
  }
  static ShardPlan shard_plan(uint32_t n_vars, uint32_t N, int rank, int world, int mode, double skew)
  {
    ShardPlan sp;
    if (mode == 0 || world < 2 || world > 64) {
      shard(N, rank, world, &sp.lo[0], &sp.hi[0]);
      uint32_t lo, hi;
      shard_skewed(n_vars, rank, world, skew, &lo, &hi);
      for (int k = 1; k < 5; ++k) {
        sp.lo[k] = lo;
        sp.hi[k] = hi;
      }
      return sp;
    }
    // fitted on B200 at 3200k constraints (tools/shard_probe.py at 2, 4 and 8 ranks): relative to a point of a fused G1
    // table, a G2 point costs 3.0 (batched affine rounds plus the longer latency-bound reduction tail of a G2 piece), an
    // H point 1.1 (its sort cannot start before the transforms end), one polynomial's iNTT + NTT 0.24 N points
    const double w2 = env_double("B200_PLAN_W2", 3.0), w_h = env_double("B200_PLAN_WH", 1.1), w_ntt = env_double("B200_PLAN_WNTT", 0.24);
    const double wt[5] = {w_h > 0.1 ? w_h : 1.1, 1.0, 1.0, 1.0, w2 > 0.1 ? w2 : 3.0};
    const uint32_t cnt[5] = {N, n_vars, n_vars, n_vars, n_vars};
    double len[5], total = 0;
    for (int k = 0; k < 5; ++k) {
      len[k] = cnt[k] * wt[k];
      total += len[k];
    }
    const double ntt = w_ntt * (double)N;
    const double target = (total + 3 * ntt) / world;
    double share[64], ssum = 0;
    for (int r = 0; r < world; ++r) {
      share[r] = target - ntt * polys_owned(r, world);
      if (share[r] < 0) share[r] = 0;
      ssum += share[r];
    }
    auto cut = [&](int r) { // position of the boundary in front of rank r on the weighted line
      if (r <= 0) return 0.0;
      if (r >= world) return total;
      double acc = 0;
      for (int q = 0; q < r; ++q)
        acc += share[q];
      double x = acc * (total / ssum), off = 0;
      for (int k = 0; k < 5; ++k) { // snap to a nearby section edge
        if (fabs(x - off) < 0.04 * len[k]) return off;
        if (fabs(x - (off + len[k])) < 0.04 * len[k]) return off + len[k];
        off += len[k];
      }
      return x;
    };
    const double b = cut(rank), e = cut(rank + 1);
    double off = 0;
    for (int k = 0; k < 5; ++k) {
      auto to_pt = [&](double x) {
        double y = (x - off) / wt[k];
        if (y <= 0) return (uint32_t)0;
        if (y >= (double)cnt[k]) return cnt[k];
        return (uint32_t)(y + 0.5);
      };
      sp.lo[k] = to_pt(b);
      sp.hi[k] = to_pt(e);
      if (sp.hi[k] < sp.lo[k]) sp.hi[k] = sp.lo[k];
      off += len[k];
    }
    return sp;
  }

  // The prover's transforms must run over the subgroup generated by host_omega(power): the coset powers (keys) and the
  // zkey's H points are tied to it.  A global domain initialised by another caller is reused only when it is large enough
  // AND its root squares down to that generator; otherwise it is replaced (bn254_ntt_init_domain accepts any primitive root).
  static eIcicleError ensure_prover_domain(uint32_t power, cudaStream_t st)
  {
    const NttDomain* d = ntt_domain();
    bool ok = d && d->max_log >= (int)power;
    if (ok) {
      Fr w = Fr::to_mont(d->root_std);
      for (int i = d->max_log; i > (int)power; --i)
        w = w.sqr();
      ok = (w == Fr::to_mont(host_omega((int)power)));
    }
    if (ok) return ICICLE_SUCCESS;
    if (d) bn254_ntt_release_domain();
    return ntt_init_domain_host(host_omega((int)power), st);
  }

  static void cache_free(b200_zkey_cache* c)
  {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {c->pH, c->row_ptr, c->col, c->val, c->keys, c->d_witness, c->d_vec, c->d_h, c->d_results, c->qx_slices, c->d_all_parts};
    for (void* p : ptrs)
      if (p) cudaFree(p);
    for (auto& g : c->groups)
    {
      for (void* p : {(void*)g.idx, (void*)g.d_w, (void*)g.g1[0], (void*)g.g1[1], (void*)g.g1[2], (void*)g.g2, (void*)g.d_out})
        if (p) cudaFree(p);
      if (g.ev_sort) cudaEventDestroy(g.ev_sort);
      if (g.ev_done) cudaEventDestroy(g.ev_done);
    }
    if (c->h_results) cudaFreeHost(c->h_results);
    if (c->h_all_parts) cudaFreeHost(c->h_all_parts);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->ev_slice) cudaEventDestroy(c->ev_slice);
    if (c->ev_xch) cudaEventDestroy(c->ev_xch);
    cudaStream_t ss[] = {c->s_copy, c->s_g1, c->s_g2, c->s_g3, c->s_q, c->s_h};
    for (auto s : ss)
      if (s) cudaStreamDestroy(s);
    cudaEvent_t es[] = {c->ev_start, c->ev_h2d, c->ev_r1cs, c->ev_ntt, c->ev_g1, c->ev_g2, c->ev_q, c->ev_prev, c->ev_b1, c->ev_free};
    for (auto e : es)
      if (e) cudaEventDestroy(e);
    delete c;
  }

  // Upload the [lo,hi) slice of the logical array  zeros(prefix) ++ section  (prefix points at infinity in front let
  // the C section, which starts at signal n_public+1, share the witness indexing of A/B1/B2); with precompute > 1
  // expand it into the [i*f + j] = 2^(shift*j) P_i table the MSM consumes (cuda_msm.cuh:29-43 layout)
  // Upload only the listed points (keep[k] = index relative to lo) of a section: the compacted B1/B2 tables
  template <class F>
  static eIcicleError upload_points_compact(
    b200_zkey_cache* c, const Section& sec, uint32_t lo, const std::vector<uint32_t>& keep, const MsmPlan& plan, Affine<F>** out,
    cudaStream_t st)
  {
    const size_t n = keep.size();
    const int f = plan.factor;
    B200_CUDA(dev_alloc(out, n * f, c), ICICLE_ALLOCATION_FAILED);
    if (n == 0) return ICICLE_SUCCESS;
    std::vector<Affine<F>> host(n);
    const Affine<F>* src = reinterpret_cast<const Affine<F>*>(sec.p) + lo;
    for (size_t k = 0; k < n; ++k)
      memcpy(&host[k], src + keep[k], sizeof(Affine<F>));
    Affine<F>* tmp = *out;
    if (f > 1) B200_CUDA(cudaMallocAsync((void**)&tmp, n * sizeof(Affine<F>), st), ICICLE_ALLOCATION_FAILED);
    B200_CUDA(cudaMemcpyAsync(tmp, host.data(), n * sizeof(Affine<F>), cudaMemcpyHostToDevice, st), ICICLE_COPY_FAILED);
    eIcicleError e = ICICLE_SUCCESS;
    if (f > 1) {
      e = precompute_enqueue<F>(tmp, true, (int)n, f, plan.c * plan.sets, *out, true, st);
      cudaFreeAsync(tmp, st);
    }
    B200_CUDA(cudaStreamSynchronize(st), ICICLE_SYNCHRONIZATION_FAILED); // `host` goes out of scope
    return e;
  }

  template <class F>
  static eIcicleError upload_points(
    b200_zkey_cache* c, const Section& sec, uint32_t prefix, uint32_t lo, uint32_t hi, const MsmPlan& plan, Affine<F>** out,
    cudaStream_t st)
  {
    const size_t n = hi - lo;
    const int f = plan.factor;
    B200_CUDA(dev_alloc(out, n * f, c), ICICLE_ALLOCATION_FAILED);
    if (n == 0) return ICICLE_SUCCESS;
    Affine<F>* tmp = *out;
    if (f > 1) B200_CUDA(cudaMallocAsync((void**)&tmp, n * sizeof(Affine<F>), st), ICICLE_ALLOCATION_FAILED);
    const size_t nz = lo < prefix ? std::min<size_t>(prefix - lo, n) : 0; // leading points at infinity in this slice
    if (nz) B200_CUDA(cudaMemsetAsync(tmp, 0, nz * sizeof(Affine<F>), st), ICICLE_COPY_FAILED);
    if (n > nz) {
      const size_t first = (lo + nz) - prefix; // index into the section
      B200_CUDA(
        cudaMemcpyAsync(tmp + nz, sec.p + first * sizeof(Affine<F>), (n - nz) * sizeof(Affine<F>), cudaMemcpyHostToDevice, st),
        ICICLE_COPY_FAILED);
    }
    if (f == 1) return ICICLE_SUCCESS;
    // the tables are built on a second stream so the next section's upload (and the host-side CSR build) overlap them
    cudaStream_t sp = c->s_q;
    B200_CUDA(cudaEventRecord(c->ev_b1, st), ICICLE_UNKNOWN_FALLBACK);
    B200_CUDA(cudaStreamWaitEvent(sp, c->ev_b1, 0), ICICLE_UNKNOWN_FALLBACK);
    eIcicleError e = precompute_enqueue<F>(tmp, true, (int)n, f, plan.c * plan.sets, *out, true, sp);
    cudaFreeAsync(tmp, sp);
    return e;
  }

  static eIcicleError cache_build(const uint8_t* zkey, size_t zkey_len, int precompute, int rank, int world, b200_zkey_cache** out)
  {
    if (!zkey || !out) return ICICLE_INVALID_POINTER;
    if (world < 1 || rank < 0 || rank >= world) return ICICLE_INVALID_ARGUMENT;
    // (format validation first: malformed input is INVALID_ARGUMENT on any machine, with or without a GPU)
    std::map<uint32_t, Section> sec;
    if (!parse_binfile(zkey, zkey_len, "zkey", 2, sec)) return ICICLE_INVALID_ARGUMENT;
    for (uint32_t id : {1u, 2u, 4u, 5u, 6u, 7u, 8u, 9u})
      if (!sec.count(id)) return ICICLE_INVALID_ARGUMENT;
    uint32_t protocol = 0;
    if (sec[1].size < 4) return ICICLE_INVALID_ARGUMENT;
    memcpy(&protocol, sec[1].p, 4);
    if (protocol != 1) return ICICLE_INVALID_ARGUMENT; // "Protocol not supported" (file_wrapper.rs:196-208)

    // header (zkey.rs:47-85)
    const Section& h = sec[2];
    const size_t need = 4 + 32 + 4 + 32 + 12 + 64 + 64 + 128 + 128 + 64 + 128;
    if (h.size < need) return ICICLE_INVALID_ARGUMENT;
    const uint8_t* p = h.p;
    uint32_t n8q, n8r;
    memcpy(&n8q, p, 4);
    if (n8q != 32 || memcmp(p + 4, FQ_MODULUS, 32) != 0) return ICICLE_INVALID_ARGUMENT; // not BN254
    memcpy(&n8r, p + 36, 4);
    if (n8r != 32 || memcmp(p + 40, FR_MODULUS, 32) != 0) return ICICLE_INVALID_ARGUMENT;
    {
      // header sanity before touching the device
      uint32_t nv, npub, dom;
      memcpy(&nv, p + 72, 4);
      memcpy(&npub, p + 76, 4);
      memcpy(&dom, p + 80, 4);
      if (dom == 0 || (dom & (dom - 1)) || nv == 0 || npub + 1 > nv) return ICICLE_INVALID_ARGUMENT;
    }
    B200_TRY(ensure_device());
    b200_zkey_cache* c = new b200_zkey_cache();
    c->device = active_device();
    c->rank = rank;
    c->world = world;
    if (precompute == 0) {
      // auto: keep the 2^(c*j) multiples of every base point resident when they fit comfortably (a third of the free
      // HBM); at 3200k that is 17.8 GB of 180 and takes the proof from ~72 ms to ~59 ms
      uint32_t nv, dom;
      memcpy(&nv, h.p + 72, 4);
      memcpy(&dom, h.p + 80, 4);
      size_t free_b = 0, total_b = 0;
      const size_t tables = ((size_t)nv * (3 * 64 + 128) + (size_t)dom * 64) * 16 / (size_t)world;
      precompute = (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && tables < free_b / 3) ? 16 : 1;
    }
    c->precompute = precompute > 1 ? precompute : 1;
    memcpy(&c->n_vars, p + 72, 4);
    memcpy(&c->n_public, p + 76, 4);
    memcpy(&c->domain_size, p + 80, 4);
    p += 84;
    memcpy(&c->alpha1, p, 64);
    memcpy(&c->beta1, p + 64, 64);
    memcpy(&c->beta2, p + 128, 128);
    /* gamma2 at p+256 is not used by the prover */
    memcpy(&c->delta1, p + 384, 64);
    memcpy(&c->delta2, p + 448, 128);
    const uint32_t N = c->domain_size;
    if (N == 0 || (N & (N - 1)) || c->n_vars == 0 || c->n_public + 1 > c->n_vars) {
      delete c;
      return ICICLE_INVALID_ARGUMENT;
    }
    while ((1u << c->power) < N)
      ++c->power;
    if (c->power + 1 > 28) {
      delete c;
      return ICICLE_INVALID_ARGUMENT;
    }
    const uint32_t n_c = c->n_vars - c->n_public - 1;
    if (sec[5].size != (uint64_t)c->n_vars * 64 || sec[6].size != (uint64_t)c->n_vars * 64 ||
        sec[7].size != (uint64_t)c->n_vars * 128 || sec[8].size != (uint64_t)n_c * 64 || sec[9].size != (uint64_t)N * 64 ||
        sec[4].size < 4) {
      delete c;
      return ICICLE_INVALID_ARGUMENT;
    }

    eIcicleError err = ICICLE_SUCCESS;
    cudaError_t ce = cudaSuccess;
    auto fail = [&](eIcicleError e) {
      cache_free(c);
      return e;
    };
#define CK(call)                                                                                                       \
  do {                                                                                                                 \
    ce = (call);                                                                                                       \
    if (ce != cudaSuccess) {                                                                                           \
      fprintf(stderr, "[icicle_b200] %s: %s\n", #call, cudaGetErrorString(ce));                                        \
      return fail(translate(ce, ICICLE_UNKNOWN_FALLBACK));                                                             \
    }                                                                                                                  \
  } while (0)

    // the quotient chain feeds the H MSM, so its (shared-memory heavy) NTT CTAs get scheduling priority over the
    // register-hungry accumulate kernels of the witness-only MSMs; G2 (the longest single MSM) comes next
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)); // lo = least urgent (numerically greatest)
    const char* pe = getenv("B200_STREAM_PRIO");
    const bool use_prio = !(pe && pe[0] == '0');
    CK(cudaStreamCreateWithPriority(&c->s_copy, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&c->s_q, cudaStreamNonBlocking, use_prio ? prio_hi : prio_lo));
    CK(cudaStreamCreateWithPriority(&c->s_g2, cudaStreamNonBlocking, use_prio && prio_hi + 1 <= prio_lo ? prio_hi + 1 : prio_lo));
    // B200_H_LAST=1 (experiment, off by default): the H MSM on its own lowest-priority stream behind the transforms, so that
    // the bucket-reduction tail left exposed at the end of the proof is H's single table rather than the three fused G1
    // tables'.  Measured 56.65 vs 56.10 ms at 3200k (profiles/r02_multi_gpu.md): H accumulating alone at the end costs
    // more than the shorter tail saves.
    const char* hl = getenv("B200_H_LAST");
    const bool h_last = use_prio && hl && hl[0] == '1' && prio_hi + 3 <= prio_lo - 1;
    const int prio_g1 = h_last ? prio_hi + 3 : prio_lo;
    CK(cudaStreamCreateWithPriority(&c->s_g1, cudaStreamNonBlocking, prio_g1));
    CK(cudaStreamCreateWithPriority(&c->s_g3, cudaStreamNonBlocking, prio_g1));
    if (h_last) CK(cudaStreamCreateWithPriority(&c->s_h, cudaStreamNonBlocking, prio_lo));
    for (cudaEvent_t* e : {&c->ev_prev, &c->ev_b1, &c->ev_free})
      CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&c->ev_start, &c->ev_h2d, &c->ev_r1cs, &c->ev_ntt, &c->ev_g1, &c->ev_g2, &c->ev_q})
      CK(cudaEventCreate(e));
    cudaStream_t st = c->s_copy;
    // B200_CACHE_TIMING=1: host-clock breakdown of the cold path on stderr
    const bool timing = getenv("B200_CACHE_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (timing)
        fprintf(stderr, "[icicle_b200] cache build: %-28s +%.3f s\n", what,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    };

    // ---- base points: this rank's part of every section (shard_plan)
    // B200_SHARD_SKEW (uniform plan; set by a caller that splits the quotient chain, see shard_skewed) and B200_SHARD_PLAN /
    // B200_PLAN_W2 / B200_PLAN_WNTT: every rank must see the same values
    const char* sk_env = getenv("B200_SHARD_SKEW");
    {
      c->plan_mode = plan_mode_default(world);
      c->plan_skew = sk_env ? atof(sk_env) : 0.0;
      ShardPlan sp = shard_plan(c->n_vars, N, rank, world, c->plan_mode, c->plan_skew);
      for (int k = 0; k < 5; ++k) {
        c->sec_lo[k] = sp.lo[k];
        c->sec_hi[k] = sp.hi[k];
      }
      c->h_lo = sp.lo[0];
      c->h_hi = sp.hi[0];
      c->w_lo = c->n_vars;
      c->w_hi = 0;
      for (int k = 1; k < 5; ++k)
        if (sp.hi[k] > sp.lo[k]) {
          c->w_lo = std::min(c->w_lo, sp.lo[k]);
          c->w_hi = std::max(c->w_hi, sp.hi[k]);
        }
      if (c->w_hi <= c->w_lo) c->w_lo = c->w_hi = 0;
    }
    const char* c_env = getenv("B200_MSM_C"); // tuning knob: window width of the cache's MSM plans (0/unset = heuristic)
    const int c_req = c_env ? atoi(c_env) : 0;
    auto plan_for = [&](uint32_t n, bool g2) {
      MsmPlan p = make_msm_plan(n ? (int)n : 1, c_req > 0 ? c_req : 0, 254, c->precompute, g2);
      p.stride = p.factor; // the cache builds its own tables: only the multiples the plan uses
      return p;
    };
    c->planH = plan_for(c->h_hi - c->h_lo, false);
    {
      // entries pack the sign in bit 31 of (point index * factor + table column); entry positions are 32-bit
      const uint64_t f = (uint64_t)c->precompute;
      if ((uint64_t)c->n_vars * f >= (1ull << 31) || (uint64_t)N * f >= (1ull << 31) || c->planH.entries() >= (1ull << 32) ||
          plan_for(c->n_vars, false).entries() >= (1ull << 32))
        return fail(ICICLE_INVALID_ARGUMENT);
    }
    {
      // which signals of a range have a B point at all? (B1 and B2 are zero together: same v_s(tau))
      std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> keep_memo; // one scan per distinct range
      auto keep_list = [&](uint32_t lo, uint32_t hi) -> const std::vector<uint32_t>& {
        auto it = keep_memo.find({lo, hi});
        if (it != keep_memo.end()) return it->second;
        std::vector<uint32_t>& keep = keep_memo[{lo, hi}];
        keep.reserve(hi - lo);
        const uint64_t* b1 = reinterpret_cast<const uint64_t*>(sec[6].p) + (size_t)lo * 8;
        const uint64_t* b2 = reinterpret_cast<const uint64_t*>(sec[7].p) + (size_t)lo * 16;
        for (uint32_t i = 0; i < hi - lo; ++i) {
          uint64_t any = 0;
          for (int k = 0; k < 8; ++k)
            any |= b1[(size_t)i * 8 + k];
          for (int k = 0; k < 16; ++k)
            any |= b2[(size_t)i * 16 + k];
          if (any) keep.push_back(i);
        }
        return keep;
      };
      const char* sp_env = getenv("B200_SPARSE_B"); // 0 = never compact, 1 = always (tests), default: >= 1/8 at infinity
      std::vector<std::vector<uint32_t>> keeps; // per group (empty for dense groups)
      // the group covering [lo, hi) (compact: with the B columns at infinity dropped), created on first use
      auto group_for = [&](uint32_t lo, uint32_t hi, bool compact) -> int {
        for (size_t g = 0; g < c->groups.size(); ++g)
          if (c->groups[g].lo == lo && c->groups[g].hi == hi && c->groups[g].compact == compact) return (int)g;
        b200_zkey_cache::WGroup g;
        g.lo = lo;
        g.hi = hi;
        g.compact = compact;
        keeps.emplace_back();
        if (compact) {
          keeps.back() = keep_list(lo, hi);
          g.n = (uint32_t)keeps.back().size();
        } else {
          g.n = hi - lo;
        }
        g.plan = plan_for(g.n, false); // B2 shares the digits/windows of the G1 tables: it reuses their sort
        c->groups.push_back(g);
        return (int)c->groups.size() - 1;
      };
      auto b_is_sparse = [&](uint32_t lo, uint32_t hi) {
        if (hi <= lo) return false;
        if (sp_env && sp_env[0] == '0') return false;
        if (sp_env && sp_env[0] == '1') return true;
        return (hi - lo - keep_list(lo, hi).size()) * 8 >= (size_t)(hi - lo);
      };
      struct SecDesc {
        int plan_idx, zkey_sec, slot; // slot: 0 = A, 1 = B1, 2 = C, -1 = B2 (G2)
        uint32_t prefix;
        bool b;
      };
      const SecDesc descs[4] = {{1, 5, 0, 0, false}, {2, 6, 1, 0, true}, {3, 8, 2, c->n_public + 1, false}, {4, 7, -1, 0, true}};
      for (const SecDesc& d : descs) {
        const uint32_t lo = c->sec_lo[d.plan_idx], hi = c->sec_hi[d.plan_idx];
        if (hi <= lo) continue;
        const bool compact = d.b && b_is_sparse(lo, hi);
        const int gi = group_for(lo, hi, compact);
        b200_zkey_cache::WGroup& g = c->groups[gi];
        if (d.slot == 1) {
          c->n_b = g.n;
          c->b_total = hi - lo;
        }
        if (compact && !g.idx && g.n) {
          CK(dev_alloc(&g.idx, (size_t)g.n, c));
          CK(dev_alloc(&g.d_w, (size_t)g.n, c));
          CK(cudaMemcpyAsync(g.idx, keeps[gi].data(), (size_t)g.n * 4, cudaMemcpyHostToDevice, st));
        }
        if (d.slot >= 0) {
          G1Affine** dst = &g.g1[g.n_g1];
          g.g1_slot[g.n_g1++] = d.slot;
          err = compact ? upload_points_compact<Fq>(c, sec[d.zkey_sec], lo, keeps[gi], g.plan, dst, st)
                        : upload_points<Fq>(c, sec[d.zkey_sec], d.prefix, lo, hi, g.plan, dst, st);
        } else {
          err = compact ? upload_points_compact<Fq2>(c, sec[d.zkey_sec], lo, keeps[gi], g.plan, &g.g2, st)
                        : upload_points<Fq2>(c, sec[d.zkey_sec], d.prefix, lo, hi, g.plan, &g.g2, st);
        }
        if (err != ICICLE_SUCCESS) return fail(err);
      }
      for (auto& g : c->groups) {
        if (g.n_g1) CK(dev_alloc(&g.d_out, 3, c));
        CK(cudaEventCreateWithFlags(&g.ev_sort, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.ev_done, cudaEventDisableTiming));
      }
      CK(cudaStreamSynchronize(st)); // the keep lists go out of scope
    }
    if ((err = upload_points<Fq>(c, sec[9], 0, c->h_lo, c->h_hi, c->planH, &c->pH, st)) != ICICLE_SUCCESS) return fail(err);
    lap("points uploaded, tables queued");

    // ---- coefficients -> CSR (record layout: cache.rs:126-166)
    const Section& cs = sec[4];
    const size_t s_coef = 12 + 32;
    uint32_t declared = 0;
    memcpy(&declared, cs.p, 4);
    c->n_coef = (cs.size - 4) / s_coef;
    // 32-bit CSR positions / entry indices: refuse what would wrap instead of reading out of bounds
    if (c->n_coef >= (1ull << 32) || declared != c->n_coef || (cs.size - 4) % s_coef != 0) return fail(ICICLE_INVALID_ARGUMENT);
    const uint8_t* rec = cs.p + 4;
    std::vector<uint32_t> row_ptr(2 * (size_t)N + 1, 0);
    for (uint64_t i = 0; i < c->n_coef; ++i) {
      const uint8_t* r = rec + i * s_coef;
      uint32_t m = r[0], row, sig;
      memcpy(&row, r + 4, 4);
      memcpy(&sig, r + 8, 4);
      if (m > 1 || row >= N || sig >= c->n_vars) return fail(ICICLE_INVALID_ARGUMENT);
      ++row_ptr[(size_t)m * N + row + 1];
    }
    for (size_t i = 0; i < 2 * (size_t)N; ++i)
      row_ptr[i + 1] += row_ptr[i];
    std::vector<uint32_t> cur(row_ptr.begin(), row_ptr.end() - 1);
    std::vector<uint32_t> col(c->n_coef ? c->n_coef : 1);
    std::vector<Fr> val(c->n_coef ? c->n_coef : 1);
    for (uint64_t i = 0; i < c->n_coef; ++i) {
      const uint8_t* r = rec + i * s_coef;
      uint32_t m = r[0], row, sig;
      memcpy(&row, r + 4, 4);
      memcpy(&sig, r + 8, 4);
      uint32_t pos = cur[(size_t)m * N + row]++;
      col[pos] = sig;
      memcpy(&val[pos], r + 12, 32);
    }
    lap("CSR built on the host");
    CK(dev_alloc(&c->row_ptr, row_ptr.size(), c));
    CK(dev_alloc(&c->col, col.size(), c));
    CK(dev_alloc(&c->val, val.size(), c));
    CK(cudaMemcpyAsync(c->row_ptr, row_ptr.data(), row_ptr.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->col, col.data(), col.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->val, val.data(), val.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));

    // ---- coset powers with 1/N folded in: keys[i] = N^-1 * w_2N^i, w_2N = W[power+1] (cache.rs:168-169,220-226)
    {
      PowTab t;
      Fr g = Fr::to_mont(host_omega((int)c->power + 1));
      for (int i = 0; i < 30; ++i) {
        t.pw[i] = g;
        g = g.sqr();
      }
      Fr two = Fr::one().dbl(), nn = Fr::one();
      for (uint32_t i = 0; i < c->power; ++i)
        nn = nn * two;
      CK(dev_alloc(&c->keys, (size_t)N, c));
      B200_LAUNCH(powers_kernel, grid_for(N, 256, 8), 256, 0, st, t, nn.inverse(), (int)c->power, c->keys);
    }

    // ---- NTT domain of order N (cache.rs:242-256 sizes it from points_a.len(); domain_size is what the
    //      transforms need - SURVEY App. C). A larger existing domain is kept; a smaller one is replaced.
    if ((err = ensure_prover_domain(c->power, st)) != ICICLE_SUCCESS) return fail(err);

    // ---- workspace
    c->wit_slice = ((size_t)c->n_vars + world - 1) / world;
    CK(dev_alloc(&c->d_witness, c->wit_slice * world, c));
    CK(dev_alloc(&c->d_vec, 3 * (size_t)N, c));
    CK(dev_alloc(&c->d_h, (size_t)N, c));
    CK(dev_alloc(&c->d_results, (size_t)4 * 96 + 192, c));
    {
      // result slots no MSM of this rank writes (sections it holds no part of) keep the identity
      static const G1Projective id1 = {Fq::zero(), Fq::raw_one(), Fq::zero()};
      static const G2Projective id2 = {Fq2::zero(), {Fq::raw_one(), Fq::zero()}, Fq2::zero()};
      for (int k = 0; k < 4; ++k)
        CK(cudaMemcpyAsync(c->d_results + 96 * k, &id1, 96, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(c->d_results + 4 * 96, &id2, 192, cudaMemcpyHostToDevice, st));
    }
    CK(cudaHostAlloc((void**)&c->h_results, 4 * 96 + 192, cudaHostAllocDefault));
    CK(cudaGetLastError());
    lap("everything queued");
    CK(cudaStreamSynchronize(st)); // host vectors go out of scope
    CK(cudaStreamSynchronize(c->s_q)); // precompute tables
    lap("device idle");
    CK(cudaEventRecord(c->ev_prev, st));
#undef CK
    *out = c;
    return ICICLE_SUCCESS;
  }

  // ---------------------------------------------------------------------------------------------- prove
  struct ResultSlots {
    G1Projective *a, *b1, *c, *h;
    G2Projective* b2;
  };
  static ResultSlots result_slots(b200_zkey_cache* c)
  {
    G1Projective* r_a = (G1Projective*)c->d_results;
    return {r_a, r_a + 1, r_a + 2, r_a + 3, (G2Projective*)(c->d_results + 4 * 96)};
  }

  // Host -> device copy of [lo, hi) of the witness on s_copy.  A pageable source (a Rust Vec, an mmap'd .wtns: what the
  // reference's callers pass) would be staged by the driver at ~12 GB/s on one thread; instead four threads copy 4 MiB
  // chunks into a pinned buffer and each chunk's DMA is queued as soon as it is complete, so the copy into pinned memory
  // overlaps the transfer.  Pinned, registered and device sources take the direct path.  B200_STAGE_MIN_BYTES (default
  // 4 MiB) is the size from which staging pays.
  static eIcicleError copy_witness_range(b200_zkey_cache* c, const bn254_scalar_t* witness, size_t lo, size_t hi)
  {
    if (hi <= lo) return ICICLE_SUCCESS;
    const size_t bytes = (hi - lo) * 32;
    const char* mb_env = getenv("B200_STAGE_MIN_BYTES");
    const size_t min_bytes = mb_env ? (size_t)atoll(mb_env) : (size_t)4 << 20;
    bool pageable = false;
    if (bytes >= min_bytes) {
      cudaPointerAttributes pa;
      pageable = cudaPointerGetAttributes(&pa, witness) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
      (void)cudaGetLastError();
    }
    if (pageable && !c->h_stage && cudaHostAlloc((void**)&c->h_stage, (size_t)c->n_vars * 32, cudaHostAllocDefault) != cudaSuccess) {
      (void)cudaGetLastError();
      c->h_stage = nullptr;
      pageable = false; // no pinned memory to be had: let the driver stage it
    }
    if (!pageable) {
      B200_CUDA(cudaMemcpyAsync(c->d_witness + lo, witness + lo, bytes, cudaMemcpyDefault, c->s_copy), ICICLE_COPY_FAILED);
      return ICICLE_SUCCESS;
    }
    const size_t chunk = (size_t)4 << 20;
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    const int n_threads = (int)std::min<size_t>(4, n_chunks);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(witness + lo);
    uint8_t* stage = c->h_stage + lo * 32;
    uint8_t* dst = reinterpret_cast<uint8_t*>(c->d_witness + lo);
    std::atomic<int> failed(0);
    auto work = [&](int t) {
      if (cudaSetDevice(c->device) != cudaSuccess) failed = 1;
      for (size_t k = (size_t)t; k < n_chunks && !failed; k += (size_t)n_threads) {
        const size_t off = k * chunk, len = std::min(chunk, bytes - off);
        memcpy(stage + off, src + off, len);
        if (cudaMemcpyAsync(dst + off, stage + off, len, cudaMemcpyHostToDevice, c->s_copy) != cudaSuccess) failed = 1;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t)
      pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool)
      th.join();
    return failed ? ICICLE_COPY_FAILED : ICICLE_SUCCESS;
  }

  // witness H2D (proof_helper.rs:194-196); every compute stream waits on it
  // (full == false: a rank that evaluates no R1CS rows only needs the witness slice its MSM shard reads)
  static eIcicleError enqueue_upload(b200_zkey_cache* c, const bn254_scalar_t* witness, uint32_t n_witness, bool full = true)
  {
    if (!c || !witness) return ICICLE_INVALID_POINTER;
    if (n_witness != c->n_vars) return ICICLE_INVALID_ARGUMENT; // "Invalid witness length" (proof_helper.rs:259-264)
    B200_CUDA(cudaSetDevice(c->device), ICICLE_INVALID_DEVICE);
    // CacheManager::get_cache re-initialises the NTT domain on every proof (cache.rs:242-256): a no-op unless
    // someone released or shrank it in between
    B200_TRY(ensure_prover_domain(c->power, c->s_copy));
    B200_CUDA(cudaEventRecord(c->ev_start, c->s_copy), ICICLE_UNKNOWN_FALLBACK);
    if (full && c->world > 1 && c->ev_slice && c->w_hi > c->w_lo) {
      // sharded rank that also evaluates R1CS rows: its own slice first - the witness MSMs (s_g1, s_g2) start on it
      // while the rest of the witness, which only the quotient chain reads, is still crossing PCIe
      B200_TRY(copy_witness_range(c, witness, c->w_lo, c->w_hi));
      B200_CUDA(cudaEventRecord(c->ev_slice, c->s_copy), ICICLE_UNKNOWN_FALLBACK);
      for (cudaStream_t s : {c->s_g1, c->s_g2, c->s_g3})
        B200_CUDA(cudaStreamWaitEvent(s, c->ev_slice, 0), ICICLE_UNKNOWN_FALLBACK);
      B200_TRY(copy_witness_range(c, witness, 0, c->w_lo));
      B200_TRY(copy_witness_range(c, witness, c->w_hi, c->n_vars));
      B200_CUDA(cudaEventRecord(c->ev_h2d, c->s_copy), ICICLE_UNKNOWN_FALLBACK);
      B200_CUDA(cudaStreamWaitEvent(c->s_q, c->ev_h2d, 0), ICICLE_UNKNOWN_FALLBACK);
      return ICICLE_SUCCESS;
    }
    const size_t w_lo = full ? 0 : c->w_lo, w_hi = full ? c->n_vars : c->w_hi;
    B200_TRY(copy_witness_range(c, witness, w_lo, w_hi));
    B200_CUDA(cudaEventRecord(c->ev_h2d, c->s_copy), ICICLE_UNKNOWN_FALLBACK);
    for (cudaStream_t s : {c->s_g1, c->s_g2, c->s_g3, c->s_q})
      B200_CUDA(cudaStreamWaitEvent(s, c->ev_h2d, 0), ICICLE_UNKNOWN_FALLBACK);
    return ICICLE_SUCCESS;
  }

  // A.w, B.w, A.w*B.w on s_q, then iNTT -> x keys -> NTT of polynomials [first, first+count) of the d_vec layout
  // (0: B.w, 1: A.w, 2: their product) into `out` (count x N elements; may be d_vec + first*N itself)
  static eIcicleError enqueue_quotient_polys(b200_zkey_cache* c, int first, int count, Fr* out)
  {
    const uint32_t N = c->domain_size;
    B200_LAUNCH(r1cs_eval_kernel, grid_for(N, 256, 8), 256, 0, c->s_q, c->row_ptr, c->col, c->val, c->d_witness, N, c->d_vec);
    cudaEventRecord(c->ev_r1cs, c->s_q);
    Fr* mine = c->d_vec + (size_t)first * N;
    B200_TRY(ntt_enqueue(mine, mine, (int)c->power, true, count, false, c->keys, c->s_q));
    return ntt_enqueue(mine, out, (int)c->power, false, count, false, nullptr, c->s_q);
  }

  // h = a.b - c over this rank's H shard (a, b, c point at the shard's first element), then the H MSM, on s_q
  static eIcicleError enqueue_h(b200_zkey_cache* c, const Fr* a, const Fr* b, const Fr* cc)
  {
    const uint32_t cnt = c->h_hi - c->h_lo;
    if (cnt) {
      B200_LAUNCH(quotient_combine_kernel, grid_for(cnt, 256, 8), 256, 0, c->s_q, a, b, cc, cnt, c->d_h + c->h_lo);
      cudaEventRecord(c->ev_ntt, c->s_q);
      cudaStream_t sh = c->s_h ? c->s_h : c->s_q;
      if (sh != c->s_q) cudaStreamWaitEvent(sh, c->ev_ntt, 0);
      B200_TRY(msm_enqueue<Fq>(c->planH, c->d_h + c->h_lo, false, c->pH, result_slots(c).h, sh));
      cudaEventRecord(c->ev_q, sh);
    } else {
      cudaEventRecord(c->ev_ntt, c->s_q);
      cudaEventRecord(c->ev_q, c->s_q);
    }
    return ICICLE_SUCCESS;
  }

  // witness-only MSMs (proof_helper.rs:198-206).  G2 parts run on s_g2, the fused G1 parts of successive groups alternate
  // between s_g1 and s_g3 so that the sort and the latency-bound bucket reduction of one group overlap the accumulation of
  // another; a group's sort runs on the stream of its G2 part when it has one (the longest consumer), else on its G1 stream
  // `gate` (optional): the accumulation phases wait for it - a polynomial owner of a sharded proof keeps the multiplier for
  // its transforms until their slices are on the wire (the other ranks' H MSMs wait for them); the sorts are not gated
  static eIcicleError enqueue_witness_msms(b200_zkey_cache* c, cudaEvent_t gate = nullptr)
  {
    ResultSlots r = result_slots(c);
    G1Projective* const slot[3] = {r.a, r.b1, r.c};
    int g1_turn = 0;
    for (const auto& g : c->groups) {
      if (!g.n) continue;
      cudaStream_t sg1 = g.n_g1 ? ((g1_turn++ & 1) ? c->s_g3 : c->s_g1) : nullptr;
      cudaStream_t ss = g.g2 ? c->s_g2 : sg1;
      const Fr* sc = c->d_witness + g.lo;
      if (g.compact) {
        B200_LAUNCH(gather_scalars_kernel, grid_for(g.n, 256, 8), 256, 0, ss, sc, g.idx, g.n, g.d_w);
        sc = g.d_w;
      }
      MsmSorted sorted;
      B200_TRY(msm_sort_enqueue(g.plan, sc, false, &sorted, ss));
      if (g.g2) {
        if (g.n_g1) {
          cudaEventRecord(g.ev_sort, ss);
          cudaStreamWaitEvent(sg1, g.ev_sort, 0);
        }
        const G2Affine* g2_tables[1] = {g.g2};
        B200_TRY(msm_reduce_enqueue<Fq2>(g.plan, sorted, g2_tables, 1, r.b2, c->s_g2, gate));
      }
      if (g.n_g1) {
        const G1Affine* g1_tables[3] = {g.g1[0], g.g1[g.n_g1 > 1 ? 1 : 0], g.g1[g.n_g1 > 2 ? 2 : 0]};
        B200_TRY(msm_reduce_enqueue<Fq>(g.plan, sorted, g1_tables, g.n_g1, g.d_out, sg1, gate));
        for (int k = 0; k < g.n_g1; ++k)
          cudaMemcpyAsync(slot[g.g1_slot[k]], g.d_out + k, 96, cudaMemcpyDeviceToDevice, sg1);
        if (g.g2) { // the sort's scratch is released after both consumers
          cudaEventRecord(g.ev_done, sg1);
          cudaStreamWaitEvent(ss, g.ev_done, 0);
        }
      }
      msm_sorted_free(&sorted, ss);
    }
    cudaEventRecord(c->ev_free, c->s_g3); // s_g3 joins s_g1
    cudaStreamWaitEvent(c->s_g1, c->ev_free, 0);
    cudaEventRecord(c->ev_g1, c->s_g1);
    cudaEventRecord(c->ev_g2, c->s_g2);
    return ICICLE_SUCCESS;
  }

  // join on s_copy, one D2H of the five partial sums
  static eIcicleError enqueue_join(b200_zkey_cache* c)
  {
    for (cudaEvent_t e : {c->ev_q, c->ev_g1, c->ev_g2})
      B200_CUDA(cudaStreamWaitEvent(c->s_copy, e, 0), ICICLE_UNKNOWN_FALLBACK);
    B200_CUDA(
      cudaMemcpyAsync(c->h_results, c->d_results, 4 * 96 + 192, cudaMemcpyDeviceToHost, c->s_copy), ICICLE_COPY_FAILED);
    B200_CUDA(cudaEventRecord(c->ev_prev, c->s_copy), ICICLE_UNKNOWN_FALLBACK);
    return ICICLE_SUCCESS;
  }

  // enqueue everything up to the device-to-host copy of the five partial sums; returns without waiting
  static eIcicleError commit_enqueue(b200_zkey_cache* c, const bn254_scalar_t* witness, uint32_t n_witness)
  {
    B200_TRY(enqueue_upload(c, witness, n_witness));
    const uint32_t N = c->domain_size;
    B200_TRY(enqueue_quotient_polys(c, 0, 3, c->d_vec)); // quotient chain + H on the (higher-priority) s_q
    B200_TRY(enqueue_h(c, c->d_vec + c->h_lo, c->d_vec + N + c->h_lo, c->d_vec + 2 * (size_t)N + c->h_lo));
    B200_TRY(enqueue_witness_msms(c));
    return enqueue_join(c);
  }

  static eIcicleError commit_wait(b200_zkey_cache* c, b200_groth16_partials* out, b200_prove_timings* tm)
  {
    if (!c || !out) return ICICLE_INVALID_POINTER;
    B200_CUDA(cudaStreamSynchronize(c->s_copy), ICICLE_SYNCHRONIZATION_FAILED);
    B200_CUDA(cudaGetLastError(), ICICLE_UNKNOWN_FALLBACK);

    // sections this rank holds no part of contribute the identity (written into the result slots at cache build)
    memcpy(out, c->h_results, 4 * 96 + 192);
    if (tm) {
      cudaEventSynchronize(c->ev_q);
      cudaEventSynchronize(c->ev_g1);
      cudaEventSynchronize(c->ev_g2);
      float t_q = 0, t_g1 = 0, t_g2 = 0;
      cudaEventElapsedTime(&tm->h2d_ms, c->ev_start, c->ev_h2d);
      cudaEventElapsedTime(&tm->r1cs_ms, c->ev_h2d, c->ev_r1cs);
      cudaEventElapsedTime(&tm->ntt_ms, c->ev_r1cs, c->ev_ntt);
      cudaEventElapsedTime(&t_q, c->ev_start, c->ev_q);
      cudaEventElapsedTime(&t_g1, c->ev_start, c->ev_g1);
      cudaEventElapsedTime(&t_g2, c->ev_start, c->ev_g2);
      tm->msm_g1_ms = t_g1;
      tm->msm_g2_ms = t_g2;
      tm->total_ms = std::max(t_q, std::max(t_g1, t_g2));
      (void)cudaGetLastError(); // timing queries must never poison the next call
    }
    return ICICLE_SUCCESS;
  }

  static eIcicleError commit_partials(
    b200_zkey_cache* c, const bn254_scalar_t* witness, uint32_t n_witness, b200_groth16_partials* out, b200_prove_timings* tm)
  {
    if (!c || !witness || !out) return ICICLE_INVALID_POINTER;
    std::lock_guard<std::mutex> g(c->mu);
    B200_TRY(commit_enqueue(c, witness, n_witness));
    return commit_wait(c, out, tm);
  }

  template <class F>
  static XYZZ<F> load_partial(const void* p)
  {
    Projective<F> pr;
    memcpy(&pr, p, sizeof(pr));
    return xyzz_from_projective(proj_to_mont(pr));
  }

  // Blinding epilogue (proof_helper.rs:274-295) on the host.  The four multiples of delta depend only on (r, s):
  // they are computed while the GPU is still busy; two scalar multiplications of MSM results remain afterwards.
  struct BlindTerms {
    Fr r, s;
    G1XYZZ r_d1, s_d1, rs_d1;
    G2XYZZ s_d2;
  };

  static eIcicleError compute_blind(const b200_zkey_cache* c, const bn254_scalar_t* r_in, const bn254_scalar_t* s_in, BlindTerms& b)
  {
    if (r_in && s_in) {
      memcpy(&b.r, r_in, 32);
      memcpy(&b.s, s_in, 32);
    } else {
      // ScalarCfg::generate_random(2) (proof_helper.rs:276-278) - but from the kernel CSPRNG (the zero-knowledge
      // property rests on r, s being unpredictable); a failed draw is an error, never a weaker generator
      if (!host_secure_random_fr(b.r) || !host_secure_random_fr(b.s)) return ICICLE_UNKNOWN_FALLBACK;
    }
    G1XYZZ d1 = G1XYZZ::from_affine(c->delta1);
    G2XYZZ d2 = G2XYZZ::from_affine(c->delta2);
    Fr rs = Fr::from_mont(Fr::to_mont(b.r) * Fr::to_mont(b.s));
    b.r_d1 = host_scalar_mul(d1, b.r);
    b.s_d1 = host_scalar_mul(d1, b.s);
    b.rs_d1 = host_scalar_mul(d1, rs);
    b.s_d2 = host_scalar_mul(d2, b.s);
    return ICICLE_SUCCESS;
  }

  static eIcicleError finish_with(
    const b200_zkey_cache* c, const b200_groth16_partials* parts, int n_parts, const BlindTerms& bt, b200_groth16_proof* proof)
  {
    if (!c || !parts || !proof) return ICICLE_INVALID_POINTER;
    if (n_parts < 1) return ICICLE_INVALID_ARGUMENT;
    const Fr &r = bt.r, &s = bt.s;
    G1XYZZ A = G1XYZZ::inf(), B1 = G1XYZZ::inf(), C = G1XYZZ::inf(), H = G1XYZZ::inf();
    G2XYZZ B2 = G2XYZZ::inf();
    for (int i = 0; i < n_parts; ++i) { // fold the per-rank partial sums
      A.add(load_partial<Fq>(&parts[i].a));
      B1.add(load_partial<Fq>(&parts[i].b1));
      C.add(load_partial<Fq>(&parts[i].c));
      H.add(load_partial<Fq>(&parts[i].h));
      B2.add(load_partial<Fq2>(&parts[i].b2));
    }
    // pi_a = A + alpha1 + r*delta1
    G1XYZZ pi_a = A;
    pi_a.madd(c->alpha1);
    pi_a.add(bt.r_d1);
    // pi_b = B2 + beta2 + s*delta2
    G2XYZZ pi_b = B2;
    pi_b.madd(c->beta2);
    pi_b.add(bt.s_d2);
    // pi_b1 = B1 + beta1 + s*delta1
    G1XYZZ pi_b1 = B1;
    pi_b1.madd(c->beta1);
    pi_b1.add(bt.s_d1);
    // pi_c = C + H + s*pi_a + r*pi_b1 - (r*s)*delta1
    G1XYZZ pi_c = C;
    pi_c.add(H);
    pi_c.add(host_double_scalar_mul(pi_a, s, pi_b1, r));
    pi_c.add(bt.rs_d1.neg());

    G1Affine a = affine_from_mont(pi_a.to_affine());
    G2Affine b = affine_from_mont(pi_b.to_affine());
    G1Affine cc = affine_from_mont(pi_c.to_affine());
    memcpy(&proof->pi_a, &a, 64);
    memcpy(&proof->pi_b, &b, 128);
    memcpy(&proof->pi_c, &cc, 64);
    return ICICLE_SUCCESS;
  }

  // ---------------------------------------------------------------------------------------------- files + JSON
  struct MappedFile {
    const uint8_t* p = nullptr;
    size_t len = 0;
    int fd = -1;
    bool open(const char* path)
    {
      fd = ::open(path, O_RDONLY);
      if (fd < 0) return false;
      struct stat sb;
      if (fstat(fd, &sb) != 0 || sb.st_size <= 0) return false;
      len = (size_t)sb.st_size;
      void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) return false;
      p = (const uint8_t*)m;
      return true;
    }
    ~MappedFile()
    {
      if (p) munmap((void*)p, len);
      if (fd >= 0) ::close(fd);
    }
  };

  // 256-bit little-endian limbs -> base-10 string (BigUint::to_str_radix(10), conversions.rs:30-39)
  static std::string to_decimal(const uint32_t limbs[8])
  {
    uint32_t v[8];
    memcpy(v, limbs, 32);
    std::string out;
    bool zero = true;
    for (int i = 0; i < 8; ++i)
      zero = zero && v[i] == 0;
    if (zero) return "0";
    char chunk[16];
    std::vector<uint32_t> parts;
    for (;;) {
      bool z = true;
      uint64_t rem = 0;
      for (int i = 7; i >= 0; --i) {
        uint64_t cur = (rem << 32) | v[i];
        v[i] = (uint32_t)(cur / 1000000000u);
        rem = cur % 1000000000u;
        z = z && v[i] == 0;
      }
      parts.push_back((uint32_t)rem);
      if (z) break;
    }
    snprintf(chunk, sizeof chunk, "%u", parts.back());
    out = chunk;
    for (size_t i = parts.size() - 1; i-- > 0;) {
      snprintf(chunk, sizeof chunk, "%09u", parts[i]);
      out += chunk;
    }
    return out;
  }

  // serde_json::to_writer_pretty of json!(Proof): alphabetical keys, 2-space indent, no trailing newline
  // (proof_helper.rs:308-316, file_wrapper.rs:105-113)
  static std::string proof_json(const b200_groth16_proof& pr)
  {
    auto fq = [&](const bn254_fq_t& x) { return "\"" + to_decimal(x.limbs) + "\""; };
    std::string s = "{\n  \"curve\": \"bn128\",\n";
    s += "  \"pi_a\": [\n    " + fq(pr.pi_a.x) + ",\n    " + fq(pr.pi_a.y) + ",\n    \"1\"\n  ],\n";
    s += "  \"pi_b\": [\n    [\n      " + fq(pr.pi_b.x.c0) + ",\n      " + fq(pr.pi_b.x.c1) + "\n    ],\n    [\n      " +
         fq(pr.pi_b.y.c0) + ",\n      " + fq(pr.pi_b.y.c1) + "\n    ],\n    [\n      \"1\",\n      \"0\"\n    ]\n  ],\n";
    s += "  \"pi_c\": [\n    " + fq(pr.pi_c.x) + ",\n    " + fq(pr.pi_c.y) + ",\n    \"1\"\n  ],\n";
    s += "  \"protocol\": \"groth16\"\n}";
    return s;
  }

  static std::string public_json(const uint8_t* witness_section, uint32_t n_public)
  {
    if (n_public == 0) return "[]";
    std::string s = "[\n";
    for (uint32_t i = 1; i <= n_public; ++i) {
      uint32_t limbs[8];
      memcpy(limbs, witness_section + (size_t)i * 32, 32);
      s += "  \"" + to_decimal(limbs) + "\"";
      s += i < n_public ? ",\n" : "\n";
    }
    s += "]";
    return s;
  }

  static bool write_file(const char* path, const std::string& s)
  {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    return fclose(f) == 0 && ok;
  }

  // process-wide CacheManager keyed "{zkey}_{device}" (src/lib.rs:44-52, cache.rs:110-114)
  static std::mutex g_cm_mu;
  static std::map<std::string, b200_zkey_cache*> g_cache_manager;

} // namespace b200

extern "C" {

eIcicleError b200_zkey_cache_create(const uint8_t* zkey, size_t zkey_len, int precompute, b200_zkey_cache** out)
{
  return cache_build(zkey, zkey_len, precompute, 0, 1, out);
}

eIcicleError
b200_zkey_cache_create_sharded(const uint8_t* zkey, size_t zkey_len, int precompute, int rank, int world, b200_zkey_cache** out)
{
  return cache_build(zkey, zkey_len, precompute, rank, world, out);
}

eIcicleError b200_zkey_cache_destroy(b200_zkey_cache* cache)
{
  cache_free(cache);
  return ICICLE_SUCCESS;
}

eIcicleError b200_zkey_cache_info(
  const b200_zkey_cache* c, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size, uint64_t* n_coef, uint64_t* device_bytes)
{
  if (!c) return ICICLE_INVALID_POINTER;
  if (n_vars) *n_vars = c->n_vars;
  if (n_public) *n_public = c->n_public;
  if (domain_size) *domain_size = c->domain_size;
  if (n_coef) *n_coef = c->n_coef;
  if (device_bytes) *device_bytes = c->device_bytes;
  return ICICLE_SUCCESS;
}

eIcicleError b200_groth16_commit_partials(
  b200_zkey_cache* cache, const bn254_scalar_t* witness, uint32_t n_witness, b200_groth16_partials* out, b200_prove_timings* tm)
{
  return commit_partials(cache, witness, n_witness, out, tm);
}

eIcicleError b200_groth16_commit_begin(
  b200_zkey_cache* c, const bn254_scalar_t* witness, uint32_t n_witness, int first_poly, int poly_count, void* out_dev)
{
  if (!c || (poly_count > 0 && !out_dev)) return ICICLE_INVALID_POINTER;
  if (first_poly < 0 || poly_count < 0 || first_poly + poly_count > 3) return ICICLE_INVALID_ARGUMENT;
  c->mu.lock();
  if (c->in_flight) { // (unreachable while `mu` is held by the pending pair; kept as a guard against misuse)
    c->mu.unlock();
    return ICICLE_INVALID_ARGUMENT;
  }
  eIcicleError e = enqueue_upload(c, witness, n_witness, poly_count > 0);
  if (e == ICICLE_SUCCESS && poly_count > 0) e = enqueue_quotient_polys(c, first_poly, poly_count, (Fr*)out_dev);
  if (e == ICICLE_SUCCESS && poly_count == 0) cudaEventRecord(c->ev_r1cs, c->s_q); // keep the phase timers well-defined
  if (e == ICICLE_SUCCESS) e = enqueue_witness_msms(c); // keep the GPU busy while the caller exchanges slices
  if (e == ICICLE_SUCCESS && cudaStreamSynchronize(c->s_q) != cudaSuccess) e = ICICLE_SYNCHRONIZATION_FAILED;
  if (e != ICICLE_SUCCESS) {
    cudaStreamSynchronize(c->s_copy); // leave no half-enqueued proof behind
    cudaStreamSynchronize(c->s_g1);
    cudaStreamSynchronize(c->s_g2);
    cudaStreamSynchronize(c->s_q);
    (void)cudaGetLastError();
    c->mu.unlock();
  } else {
    c->in_flight = true; // `mu` stays held until commit_end
  }
  return e;
}

eIcicleError b200_groth16_commit_end(
  b200_zkey_cache* c, const void* a_dev, const void* b_dev, const void* c_dev, b200_groth16_partials* out, b200_prove_timings* tm)
{
  if (!c) return ICICLE_INVALID_POINTER;
  if (!c->in_flight) return ICICLE_INVALID_ARGUMENT; // no successful commit_begin is pending on this cache
  eIcicleError e = (!a_dev || !b_dev || !c_dev || !out) ? ICICLE_INVALID_POINTER : ICICLE_SUCCESS;
  if (e == ICICLE_SUCCESS) e = enqueue_h(c, (const Fr*)a_dev, (const Fr*)b_dev, (const Fr*)c_dev);
  if (e == ICICLE_SUCCESS) e = enqueue_join(c);
  if (e == ICICLE_SUCCESS) e = commit_wait(c, out, tm);
  if (e != ICICLE_SUCCESS) { // drain whatever commit_begin enqueued so the next proof starts clean
    for (cudaStream_t st : {c->s_copy, c->s_g1, c->s_g2, c->s_q, c->s_h})
      if (st) cudaStreamSynchronize(st);
    (void)cudaGetLastError();
  }
  c->in_flight = false;
  c->mu.unlock();
  return e;
}

eIcicleError b200_zkey_cache_b_points(const b200_zkey_cache* c, uint32_t* kept, uint32_t* total)
{
  if (!c || !kept || !total) return ICICLE_INVALID_POINTER;
  *kept = c->n_b;
  *total = c->b_total;
  return ICICLE_SUCCESS;
}

// The cut of the five base-point sections (0 = H, 1 = A, 2 = B1, 3 = C, 4 = B2) a rank of `world` holds.
// mode: 0 = uniform (every section cut `world` ways; `skew` as in b200_shard_range), 1 = line (sections laid end to end by
// cost and cut into equal-cost pieces), -1 = the library default (B200_SHARD_PLAN; line for world > 1).
eIcicleError b200_shard_plan(uint32_t n_vars, uint32_t domain_size, int rank, int world, int mode, double skew, uint32_t* lo5, uint32_t* hi5)
{
  if (!lo5 || !hi5) return ICICLE_INVALID_POINTER;
  if (world < 1 || rank < 0 || rank >= world || mode < -1 || mode > 1) return ICICLE_INVALID_ARGUMENT;
  const ShardPlan sp = shard_plan(n_vars, domain_size, rank, world, mode < 0 ? plan_mode_default(world) : mode, skew);
  for (int k = 0; k < 5; ++k) {
    lo5[k] = sp.lo[k];
    hi5[k] = sp.hi[k];
  }
  return ICICLE_SUCCESS;
}

// the mode b200_zkey_cache_create_sharded cuts with for this world size (environment included): 0 uniform, 1 line
int b200_shard_plan_mode(int world) { return plan_mode_default(world); }

// the ranges this cache was built with
eIcicleError b200_zkey_cache_ranges(const b200_zkey_cache* c, uint32_t* lo5, uint32_t* hi5)
{
  if (!c || !lo5 || !hi5) return ICICLE_INVALID_POINTER;
  for (int k = 0; k < 5; ++k) {
    lo5[k] = c->sec_lo[k];
    hi5[k] = c->sec_hi[k];
  }
  return ICICLE_SUCCESS;
}

eIcicleError b200_shard_range(uint32_t n, int rank, int world, double skew, uint32_t* lo, uint32_t* hi)
{
  if (!lo || !hi) return ICICLE_INVALID_POINTER;
  if (world < 1 || rank < 0 || rank >= world) return ICICLE_INVALID_ARGUMENT;
  shard_skewed(n, rank, world, skew, lo, hi);
  return ICICLE_SUCCESS;
}

eIcicleError b200_zkey_cache_h_range(const b200_zkey_cache* c, uint32_t* lo, uint32_t* hi)
{
  if (!c || !lo || !hi) return ICICLE_INVALID_POINTER;
  *lo = c->h_lo;
  *hi = c->h_hi;
  return ICICLE_SUCCESS;
}

eIcicleError b200_groth16_finish(
  const b200_zkey_cache* cache, const b200_groth16_partials* parts, int n_parts, const bn254_scalar_t* r,
  const bn254_scalar_t* s, b200_groth16_proof* proof)
{
  if (!cache) return ICICLE_INVALID_POINTER;
  BlindTerms bt;
  B200_TRY(compute_blind(cache, r, s, bt));
  return finish_with(cache, parts, n_parts, bt, proof);
}

eIcicleError b200_groth16_prove(
  b200_zkey_cache* cache, const bn254_scalar_t* witness, uint32_t n_witness, const bn254_scalar_t* r, const bn254_scalar_t* s,
  b200_groth16_proof* proof, b200_prove_timings* tm)
{
  if (!cache || !proof) return ICICLE_INVALID_POINTER;
  if (cache->world != 1) return ICICLE_INVALID_ARGUMENT; // sharded caches go through commit_partials + finish
  b200_groth16_partials parts;
  std::lock_guard<std::mutex> g(cache->mu);
  B200_TRY(commit_enqueue(cache, witness, n_witness));
  BlindTerms bt;
  eIcicleError be = compute_blind(cache, r, s, bt); // host work overlapped with the GPU
  B200_TRY(commit_wait(cache, &parts, tm));       // (always drain the enqueued proof before returning)
  B200_TRY(be);
  return finish_with(cache, &parts, 1, bt, proof);
}

// One proof over `comm->world` GPUs, one process per GPU, every rank calling with the same witness (HOST or DEVICE
// memory) and its own shard of the cache (b200_zkey_cache_create_sharded with the communicator's rank / world).  The whole
// data plane is inside the library: the quotient chain is split by polynomial (rank j transforms polynomial j),
// the owners send every rank its H-shard slice with ONE grouped ncclSend/ncclRecv on the chain's stream, the H MSM follows
// on the same stream, the witness MSMs run meanwhile on their own streams, one 576 B ncclAllGather collects the partial
// sums and rank 0 folds + blinds.  The host waits once, at the end.  `proof` is written on rank 0 only (may be NULL elsewhere).
eIcicleError b200_groth16_prove_sharded(
  b200_zkey_cache* c, b200_comm* comm, const bn254_scalar_t* witness, uint32_t n_witness, const bn254_scalar_t* r,
  const bn254_scalar_t* s, b200_groth16_proof* proof, b200_prove_timings* tm)
{
  if (!c || !comm || !witness) return ICICLE_INVALID_POINTER;
  if (comm->rank != c->rank || comm->world != c->world || (comm->rank == 0 && !proof)) return ICICLE_INVALID_ARGUMENT;
  if (!nccl().ok) return ICICLE_API_NOT_IMPLEMENTED;
  const int world = c->world, rank = c->rank;
  const uint32_t N = c->domain_size, cnt = c->h_hi - c->h_lo;
  auto owner_of = [world](int j) { return world >= 3 ? j : (world == 2 ? (j < 2 ? 0 : 1) : 0); };
  int first = 0, count = 0;
  for (int j = 2; j >= 0; --j)
    if (owner_of(j) == rank) {
      first = j;
      ++count;
    }
  std::lock_guard<std::mutex> g(c->mu);
  B200_CUDA(cudaSetDevice(c->device), ICICLE_INVALID_DEVICE);
  if (!c->d_all_parts) { // first sharded proof on this cache: exchange buffers
    B200_CUDA(dev_alloc(&c->qx_slices, 3 * (size_t)(cnt ? cnt : 1), c), ICICLE_ALLOCATION_FAILED);
    B200_CUDA(dev_alloc(&c->d_all_parts, (size_t)world * sizeof(b200_groth16_partials), c), ICICLE_ALLOCATION_FAILED);
    B200_CUDA(cudaMallocHost((void**)&c->h_all_parts, (size_t)world * sizeof(b200_groth16_partials)), ICICLE_ALLOCATION_FAILED);
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_slice, cudaEventDisableTiming), ICICLE_UNKNOWN_FALLBACK);
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_xch, cudaEventDisableTiming), ICICLE_UNKNOWN_FALLBACK);
  }
  // A HOST witness crosses PCIe once per box, not once per GPU: every rank uploads 1/world of it and one ncclAllGather
  // over NVLink completes it everywhere (8 GPUs pulling 102 MB each through the host cost 6.7 ms at 3200k; this is 0.6 ms).
  // A DEVICE witness is copied as before.
  eIcicleError e = ICICLE_SUCCESS;
  bool wit_on_host = true;
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, witness) == cudaSuccess)
      wit_on_host = !(pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
  }
  if (world > 1 && wit_on_host && c->wit_slice > 0) { // (every rank must pass the same kind of memory: this is a collective)
    if (n_witness != c->n_vars) return ICICLE_INVALID_ARGUMENT;
    const size_t sl = c->wit_slice, lo = std::min((size_t)rank * sl, (size_t)c->n_vars), hi = std::min(lo + sl, (size_t)c->n_vars);
    B200_TRY(ensure_prover_domain(c->power, c->s_copy));
    cudaEventRecord(c->ev_start, c->s_copy);
    e = copy_witness_range(c, witness, lo, hi);
    if (e == ICICLE_SUCCESS &&
        nccl().AllGather(c->d_witness + (size_t)rank * sl, c->d_witness, sl * 32, ncclUint8, comm->comm, c->s_copy) != ncclSuccess)
      e = (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
    cudaEventRecord(c->ev_h2d, c->s_copy);
    for (cudaStream_t st : {c->s_g1, c->s_g2, c->s_g3, c->s_q})
      cudaStreamWaitEvent(st, c->ev_h2d, 0);
  } else {
    e = enqueue_upload(c, witness, n_witness, count > 0);
  }
  if (e == ICICLE_SUCCESS && count > 0) e = enqueue_quotient_polys(c, first, count, c->d_vec + (size_t)first * N);
  if (e == ICICLE_SUCCESS && count == 0) cudaEventRecord(c->ev_r1cs, c->s_q);
  // ranks that own no polynomial start their witness MSMs at once; the owners queue theirs behind the exchange (below)
  if (e == ICICLE_SUCCESS && count == 0) e = enqueue_witness_msms(c);
  if (e == ICICLE_SUCCESS) {
    // the exchange: polynomial j's owner holds all of it; rank q needs the range of its H piece (shard_plan)
    ncclResult_t nr = nccl().GroupStart();
    for (int j = 0; j < 3 && nr == ncclSuccess; ++j) {
      const int owner = owner_of(j);
      Fr* mine = c->qx_slices + (size_t)j * cnt;
      if (owner == rank) {
        const Fr* poly = c->d_vec + (size_t)j * N;
        for (int q = 0; q < world && nr == ncclSuccess; ++q) {
          const ShardPlan spq = shard_plan(c->n_vars, N, q, world, c->plan_mode, c->plan_skew);
          const uint32_t lo = spq.lo[0], hi = spq.hi[0];
          if (hi == lo) continue;
          if (q == rank)
            cudaMemcpyAsync(mine, poly + lo, (size_t)(hi - lo) * 32, cudaMemcpyDeviceToDevice, c->s_q);
          else
            nr = nccl().Send(poly + lo, (size_t)(hi - lo) * 32, ncclUint8, q, comm->comm, c->s_q);
        }
      } else if (cnt) {
        nr = nccl().Recv(mine, (size_t)cnt * 32, ncclUint8, owner, comm->comm, c->s_q);
      }
    }
    ncclResult_t ge = nccl().GroupEnd();
    if (nr != ncclSuccess || ge != ncclSuccess) {
      fprintf(stderr, "[icicle_b200] quotient exchange: %s\n", nccl().GetErrorString(nr != ncclSuccess ? nr : ge));
      e = (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
    }
  }
  if (e == ICICLE_SUCCESS && count > 0) {
    cudaEventRecord(c->ev_xch, c->s_q);
    e = enqueue_witness_msms(c, c->ev_xch);
  }
  // d_vec order: 0 = B.w', 1 = A.w', 2 = product'; enqueue_h takes (a, b, c) = (A', B', product')
  if (e == ICICLE_SUCCESS) e = enqueue_h(c, c->qx_slices + (size_t)cnt, c->qx_slices, c->qx_slices + 2 * (size_t)cnt);
  BlindTerms bt;
  eIcicleError be = ICICLE_SUCCESS;
  if (e == ICICLE_SUCCESS) {
    for (cudaEvent_t ev : {c->ev_q, c->ev_g1, c->ev_g2})
      cudaStreamWaitEvent(c->s_copy, ev, 0);
    ncclResult_t nr = nccl().AllGather(c->d_results, c->d_all_parts, sizeof(b200_groth16_partials), ncclUint8, comm->comm, c->s_copy);
    if (nr != ncclSuccess) e = (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
    if (rank == 0)
      cudaMemcpyAsync(c->h_all_parts, c->d_all_parts, (size_t)world * sizeof(b200_groth16_partials), cudaMemcpyDeviceToHost, c->s_copy);
    cudaEventRecord(c->ev_prev, c->s_copy);
    if (rank == 0) be = compute_blind(c, r, s, bt); // host work overlapped with the GPU
  }
  // the one host wait of the proof (also drains a failed enqueue so the next proof starts clean)
  for (cudaStream_t st : {c->s_q, c->s_h, c->s_g1, c->s_g2, c->s_copy})
    if (st && cudaStreamSynchronize(st) != cudaSuccess && e == ICICLE_SUCCESS) e = ICICLE_SYNCHRONIZATION_FAILED;
  if (cudaGetLastError() != cudaSuccess && e == ICICLE_SUCCESS) e = (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
  if (e != ICICLE_SUCCESS) return e;
  if (tm) {
    float t_q = 0, t_g1 = 0, t_g2 = 0;
    cudaEventElapsedTime(&tm->h2d_ms, c->ev_start, c->ev_h2d);
    cudaEventElapsedTime(&tm->r1cs_ms, c->ev_h2d, c->ev_r1cs);
    cudaEventElapsedTime(&tm->ntt_ms, c->ev_r1cs, c->ev_ntt);
    cudaEventElapsedTime(&t_q, c->ev_start, c->ev_q);
    cudaEventElapsedTime(&t_g1, c->ev_start, c->ev_g1);
    cudaEventElapsedTime(&t_g2, c->ev_start, c->ev_g2);
    tm->msm_g1_ms = t_g1;
    tm->msm_g2_ms = t_g2;
    tm->total_ms = std::max(t_q, std::max(t_g1, t_g2));
    (void)cudaGetLastError();
  }
  if (rank != 0) return ICICLE_SUCCESS;
  B200_TRY(be);
  return finish_with(c, (const b200_groth16_partials*)c->h_all_parts, world, bt, proof);
}

// proof.json text for a proof struct (what b200_groth16_prove_files writes); returns the length, or 0 if cap is too small
size_t b200_proof_to_json(const b200_groth16_proof* proof, char* out, size_t cap)
{
  if (!proof || !out) return 0;
  std::string s = proof_json(*proof);
  if (s.size() + 1 > cap) return 0;
  memcpy(out, s.c_str(), s.size() + 1);
  return s.size();
}

eIcicleError b200_groth16_prove_files(
  const char* witness_path, const char* zkey_path, const char* proof_path, const char* public_path, const char* device)
{
  if (!witness_path || !zkey_path || !proof_path || !public_path || !device) return ICICLE_INVALID_POINTER;
  if (strncmp(device, "CUDA", 4) != 0) return ICICLE_INVALID_DEVICE; // no CPU backend behind this library
  icicleDevice dev;
  memset(&dev, 0, sizeof dev);
  strncpy(dev.type, "CUDA", sizeof dev.type - 1);
  dev.id = 0; // src/lib.rs:29
  B200_TRY(icicle_set_device(&dev));

  b200_zkey_cache* cache = nullptr;
  {
    std::lock_guard<std::mutex> g(g_cm_mu);
    std::string key = std::string(zkey_path) + "_" + device;
    auto it = g_cache_manager.find(key);
    if (it == g_cache_manager.end()) {
      MappedFile zk;
      if (!zk.open(zkey_path)) return ICICLE_INVALID_ARGUMENT;
      const char* pf = getenv("B200_PRECOMPUTE");
      B200_TRY(cache_build(zk.p, zk.len, pf ? atoi(pf) : 0 /* auto */, 0, 1, &cache));
      g_cache_manager[key] = cache;
    } else {
      cache = it->second;
    }
  }
  MappedFile wt;
  if (!wt.open(witness_path)) return ICICLE_INVALID_ARGUMENT;
  std::map<uint32_t, Section> sec;
  if (!parse_binfile(wt.p, wt.len, "wtns", 2, sec) || !sec.count(1) || !sec.count(2)) return ICICLE_INVALID_ARGUMENT;
  // header: u32 n8, n8 bytes prime, u32 n_witness (file_wrapper.rs:169-177)
  if (sec[1].size < 40) return ICICLE_INVALID_ARGUMENT;
  uint32_t n8, n_witness;
  memcpy(&n8, sec[1].p, 4);
  if (n8 != 32 || memcmp(sec[1].p + 4, FR_MODULUS, 32) != 0) return ICICLE_INVALID_ARGUMENT; // curve mismatch
  memcpy(&n_witness, sec[1].p + 36, 4);
  if (sec[2].size < (uint64_t)n_witness * 32) return ICICLE_INVALID_ARGUMENT;
  b200_groth16_proof proof;
  // the reference's `no-randomness` cargo feature (r = s = 1: proofs are NOT zero-knowledge); honoured only for the
  // exact value "1" and announced, so that a stray variable cannot silently disable blinding in production
  const char* nr_env = getenv("B200_NO_RANDOMNESS");
  const bool fixed = nr_env && strcmp(nr_env, "1") == 0;
  if (fixed) fprintf(stderr, "[icicle_b200] WARNING: B200_NO_RANDOMNESS=1 - blinding disabled (r = s = 1), proofs are not zero-knowledge\n");
  bn254_scalar_t one;
  memset(&one, 0, sizeof one);
  one.limbs[0] = 1;
  B200_TRY(b200_groth16_prove(
    cache, (const bn254_scalar_t*)sec[2].p, n_witness, fixed ? &one : nullptr, fixed ? &one : nullptr, &proof, nullptr));
  if (!write_file(proof_path, proof_json(proof))) return ICICLE_INVALID_ARGUMENT;
  if (!write_file(public_path, public_json(sec[2].p, cache->n_public))) return ICICLE_INVALID_ARGUMENT;
  return ICICLE_SUCCESS;
}

} // extern "C"
