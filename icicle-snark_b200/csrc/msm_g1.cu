// BN254 G1 MSM instantiation + plan selection + the bn254_msm / bn254_msm_precompute_bases symbols
// (/root/reference/icicle/src/msm.cpp:12-16,45-49).
#include <cstring>
#include <mutex>
#include <vector>

#include "msm_impl.cuh"

namespace b200 {

  unsigned long long g_launches = 0;
  int g_profile_mode = 0;
  static cudaEvent_t g_profile_events[2] = {nullptr, nullptr};
  static std::vector<MsmProfileRec> g_profile_recs;
  static std::mutex g_profile_mu;

  void msm_profile_begin(cudaStream_t st)
  {
    if (!g_profile_mode) return;
    cudaDeviceSynchronize();
    if (!g_profile_events[0]) {
      cudaEventCreate(&g_profile_events[0]);
      cudaEventCreate(&g_profile_events[1]);
    }
    cudaEventRecord(g_profile_events[0], st);
  }
  void msm_profile_end(cudaStream_t st, const MsmPlan& plan, int g2, int nsel, int batched)
  {
    if (!g_profile_mode || !g_profile_events[0]) return;
    cudaEventRecord(g_profile_events[1], st);
    cudaDeviceSynchronize();
    float ms = -1.f;
    cudaEventElapsedTime(&ms, g_profile_events[0], g_profile_events[1]);
    std::lock_guard<std::mutex> g(g_profile_mu);
    g_profile_recs.push_back({g2, nsel, plan.n, plan.windows, plan.c, plan.factor, plan.nbuckets, batched, ms});
  }

  MsmPlan make_msm_plan(int n, int c_req, int bitsize, int factor, bool g2)
  {
    MsmPlan p;
    p.n = n;
    int c = c_req;
    if (c <= 0) {
      // window width: ~log2(n) - 5 keeps accumulate work (n * windows adds) well above the
      // bucket-reduction work (2 * sets * 2^(c-1) full adds); the reference uses log2(n) - 4
      // with unsigned digits (cuda_msm.cuh:45-48), i.e. the same bucket count.
      int lg = 0;
      while ((1ll << lg) < n) ++lg;
      c = lg - 5;
      if (g2) c -= 1;
      // every window folded into one bucket set by precomputed bases: the reduction is 1/W of the work,
      // so wider windows (fewer adds per scalar) pay
      if (c < 2) c = 2;
      int narrow = 4;
      if (factor > 1 && lg >= 12) {
        // precomputed multiples fold the windows into few bucket sets, so the reduction is cheap and wider windows
        // (fewer adds per scalar) pay. Prefer a width whose window count fits the factor (a single bucket set).
        // Measured on B200 (tools/sweep_small.sh, tools/probe_shard.py, factor 16): 2^20 points and above want
        // c = 20; below that c = 17 wins by 5-30 % (16 windows, the top one holding only the recoding carry)
        // because the 2^19-bucket reduction stops amortising. Widths whose top window holds 1..8 bits (18, 19, 21)
        // lose everywhere: n / 2^bits entries land in each of a few buckets.
        int c_fit = (bitsize + 2 + factor - 1) / factor; // smallest c with windows <= factor
        int c_wide = lg >= 20 ? (lg + 1 < 21 ? lg + 1 : 21) : 17;
        if (c_fit <= 21) {
          c = c_wide > c_fit ? c_wide : c_fit;
          narrow = 8;
        }
      }
      // a top window holding only a few bits of the scalar funnels n/2^bits entries into each of a few
      // buckets; step down until the top window is either empty or reasonably wide
      while (c > 2) {
        int w = (bitsize + 2 + c - 1) / c;
        int top_bits = bitsize - c * (w - 1);
        if (top_bits > 0 && top_bits <= narrow) --c; else break;
      }
    }
    if (c < 2) c = 2;
    if (c > 22) c = 22;
    p.c = c;
    p.windows = (bitsize + 2 + c - 1) / c;
    p.factor = factor < 1 ? 1 : factor;
    p.stride = p.factor; // the table keeps the caller's layout out[i*f + j] even when fewer multiples are used
    if (p.factor > p.windows) p.factor = p.windows;
    p.sets = (p.windows + p.factor - 1) / p.factor;
    p.bpw = 1 << (c - 1);
    p.nbuckets = p.sets * p.bpw;
    p.item_cap = 256;
    // H = sum_w 2^(c-1) * 2^(c*w)
    for (int i = 0; i < 9; ++i)
      p.hconst[i] = 0;
    for (int w = 0; w < p.windows; ++w) {
      int bit = w * c + c - 1;
      p.hconst[bit >> 5] |= 1u << (bit & 31);
    }
    return p;
  }

  template eIcicleError msm_enqueue<Fq>(const MsmPlan&, const Fr*, bool, const Affine<Fq>*, Projective<Fq>*, cudaStream_t);
  template eIcicleError msm_reduce_enqueue<Fq>(const MsmPlan&, const MsmSorted&, const Affine<Fq>* const*, int, Projective<Fq>*, cudaStream_t, cudaEvent_t);
  template eIcicleError precompute_enqueue<Fq>(const Affine<Fq>*, bool, int, int, int, Affine<Fq>*, bool, cudaStream_t);

} // namespace b200

using namespace b200;

extern "C" {

// host-only: the plan the MSM entry points and the ZKeyCache derive for a shape (window width heuristic, window count,
// bucket sets after precompute folding, recoding constant); lets the CPU tests check the host logic without a GPU
eIcicleError b200_msm_plan_info(int n, int c, int bitsize, int precompute_factor, int g2, int32_t* out8, uint32_t* hconst9)
{
  if (!out8) return ICICLE_INVALID_POINTER;
  if (n < 1 || bitsize < 1 || bitsize > 254 || c < 0 || c > 22) return ICICLE_INVALID_ARGUMENT;
  MsmPlan p = make_msm_plan(n, c, bitsize, precompute_factor, g2 != 0);
  const int32_t v[8] = {p.c, p.windows, p.factor, p.sets, p.bpw, p.nbuckets, p.item_cap, p.n};
  for (int i = 0; i < 8; ++i)
    out8[i] = v[i];
  if (hconst9)
    for (int i = 0; i < 9; ++i)
      hconst9[i] = p.hconst[i];
  return ICICLE_SUCCESS;
}

eIcicleError bn254_msm(
  const bn254_scalar_t* scalars, const bn254_affine_t* bases, int msm_size, const MSMConfig* config, bn254_projective_t* results)
{
  return msm_api<Fq>(scalars, bases, msm_size, config, results, false);
}

eIcicleError bn254_msm_precompute_bases(
  const bn254_affine_t* input_bases, int bases_size, const MSMConfig* config, bn254_affine_t* output_bases)
{
  return precompute_api<Fq>(input_bases, bases_size, config, output_bases, false);
}

unsigned long long b200_launch_count(void) { return g_launches; }

// enable != 0: profiling mode on (see msm.cuh: every MSM's bucket-accumulation phase runs isolated and is timed),
// records cleared; enable == 0: off.  Returns the duration in ms of the last recorded accumulation phase, or -1.
float b200_profile_accumulate(int enable)
{
  std::lock_guard<std::mutex> g(g_profile_mu);
  float ms = g_profile_recs.empty() ? -1.f : g_profile_recs.back().ms;
  if (enable && !g_profile_mode) g_profile_recs.clear();
  g_profile_mode = enable ? 1 : 0;
  return ms;
}

// copies up to `cap` records {g2, nsel, n, windows, c, factor, nbuckets, batched_rounds, ms as float bits} (9 x 32-bit
// words each) collected since profiling was enabled; returns the number of records available
int b200_profile_records(int32_t* out9, int cap)
{
  std::lock_guard<std::mutex> g(g_profile_mu);
  const int n = (int)g_profile_recs.size();
  for (int i = 0; i < n && i < cap && out9; ++i) {
    const MsmProfileRec& r = g_profile_recs[i];
    const int32_t v[8] = {r.g2, r.nsel, r.n, r.windows, r.c, r.factor, r.nbuckets, r.batched};
    memcpy(out9 + 9 * i, v, sizeof v);
    memcpy(out9 + 9 * i + 8, &r.ms, 4);
  }
  return n;
}

} // extern "C"
