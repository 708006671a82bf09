// HOST model of bucket accumulation by batched affine addition - the round structure the
// round-2 kernels will have, executed serially and checked against the XYZZ accumulation the MSM uses today.
//
//   round r:  every bucket segment of L elements becomes ceil(L/2): slot j <- in[2j] + in[2j+1], an odd tail is copied
//   pass 1:   "thread" t owns BA_M consecutive slots: classifies each pair, stores the running product of the
//             denominators in front of it (32 B per slot) and its chunk total            (ba_prefix_thread)
//   pass 2:   the chunk totals are inverted with a second level of the same trick, one true inversion per BA_M2
//             chunks                                                                      (ba_invert_thread)
//   pass 3:   thread t walks its slots backwards: inverse of its denominator = prefix * running inverse, finishes
//             the add                                                                     (ba_finish_thread)
// The passes are the per-thread bodies of batch_affine.cuh - the same functions the kernels of msm_batch_affine.cuh
// wrap - called for one "thread" after another.
//
// Field products per add: 1 (prefix) + 2 (back-substitution) + 3 (chord) + (3 + I/BA_M2)/BA_M for the second level,
// I = one inversion (384 products as Fp::inverse does it): 6.9 at BA_M = 16, BA_M2 = 32.  Today's mixed XYZZ add costs 10.
// Test/model code: exported as b200_batch_affine_selfcheck for tests/test_host_math.py.
#include <algorithm>
#include <random>
#include <vector>

#include "batch_affine.cuh"
#include "common.cuh"
#include "host_math.h"

namespace b200 {

  struct ModelCounts {
    uint64_t adds = 0, products = 0, inversions = 0;
  };

  // runs the passes of a round serially: for every selection, `threads` threads (the grid bound the device launcher
  // computes), each executing the per-thread body; flags a bound that would not cover the true slot count
  template <class F>
  struct BaHostExec {
    int overflow = 0;
    ModelCounts* cnt;
    void next_offsets(const uint32_t* off, int nb, uint32_t* off_next)
    {
      off_next[0] = 0;
      for (int b = 0; b < nb; ++b)
        off_next[b + 1] = off_next[b] + (off[b + 1] - off[b] + 1) / 2;
    }
    void prefix(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      const uint32_t S = L.off_next[L.nb];
      if (S > L.pts_stride || ba_threads_for(S) > threads || ba_threads_for(S) > L.tot_stride) ++overflow;
      for (int which = 0; which < nsel; ++which) {
        BaRound<F> R = ba_round_of(L, which);
        for (size_t t = 0; t < threads; ++t)
          ba_prefix_thread(R, (uint32_t)t);
      }
      // accounting (first selection): real adds of this round and the products the three passes spend on them
      BaRound<F> R = ba_round_of(L, 0);
      for (uint32_t sl = 0; sl < S; ++sl) {
        int b = ba_find_bucket(L.off_next, L.nb, sl);
        Affine<F> a, bp;
        F den;
        bool pair = ba_operands(R, b, sl, a, bp);
        int kind = pair_prepare(a, bp, den);
        if (pair && kind <= PAIR_TANGENT) {
          ++cnt->adds;
          cnt->products += kind == PAIR_CHORD ? 3 : 4;
        }
      }
      const uint64_t T = ba_threads_for(S), U = (T + BA_M2 - 1) / BA_M2;
      cnt->products += 3 * (uint64_t)S + 3 * T + 384 * U;
      cnt->inversions += U;
    }
    void invert(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      const uint32_t n_totals = (uint32_t)ba_threads_for(L.off_next[L.nb]);
      if ((n_totals + BA_M2 - 1) / BA_M2 > threads) ++overflow;
      for (int which = 0; which < nsel; ++which) {
        BaRound<F> R = ba_round_of(L, which);
        for (size_t u = 0; u < threads; ++u)
          ba_invert_thread(R.totals, n_totals, (uint32_t)u);
      }
    }
    void finish(const BaLaunch<F>& L, size_t threads, int nsel)
    {
      for (int which = 0; which < nsel; ++which) {
        BaRound<F> R = ba_round_of(L, which);
        for (size_t t = threads; t-- > 0;)
          ba_finish_thread(R, (uint32_t)t);
      }
    }
  };

  template <class F>
  static int run_batch_affine_model(const Affine<F>& gen, int n_entries, int nb, int max_rounds, uint32_t seed, ModelCounts& cnt)
  {
    std::mt19937 rng(seed);
    // two base-point tables sharing one sort (as A/B1/C do): small multiples of the generator, so equal points,
    // opposite points and the identity all occur; the second table is the first one reversed and negated
    const int n_pts = 24, nsel = 2;
    std::vector<Affine<F>> table(n_pts), table2(n_pts);
    {
      XYZZ<F> g = XYZZ<F>::from_affine(gen), acc = XYZZ<F>::inf();
      for (int i = 0; i < n_pts; ++i) {
        if (i % 8 != 7) acc.add(g);                       // every 8th point repeats its predecessor
        table[i] = (i % 11 == 10) ? Affine<F>::inf() : acc.to_affine(); // a few identities (zkeys do contain (0,0))
      }
      for (int i = 0; i < n_pts; ++i)
        table2[i] = table[n_pts - 1 - i].is_inf() ? Affine<F>::inf() : table[n_pts - 1 - i].neg();
    }
    const Affine<F>* tables[BA_MAX_SEL] = {table.data(), table2.data(), nullptr, nullptr};
    // entries (point index | sign << 31), bucket-sorted, with offsets - the MSM's sort output
    std::vector<uint32_t> bucket_of(n_entries), offsets(nb + 1, 0);
    for (int k = 0; k < n_entries; ++k) {
      // skewed: a quarter of the entries land in bucket 0, some buckets stay empty
      uint32_t b = (rng() & 3) == 0 ? 0 : rng() % (uint32_t)nb;
      if (nb > 4 && b == 3) b = 2;
      bucket_of[k] = b;
      ++offsets[b + 1];
    }
    uint32_t max_len = 0;
    for (int b = 0; b < nb; ++b) {
      max_len = std::max(max_len, offsets[b + 1]);
      offsets[b + 1] += offsets[b];
    }
    std::vector<uint32_t> entries(n_entries ? n_entries : 1), cursor(offsets.begin(), offsets.end() - 1);
    for (int k = 0; k < n_entries; ++k)
      entries[cursor[bucket_of[k]]++] = (rng() % n_pts) | ((rng() & 1u) << 31);

    // reference: what msm_accumulate_kernel computes (XYZZ += affine)
    auto reference = [&](const Affine<F>* tab, int b) {
      XYZZ<F> acc = XYZZ<F>::inf();
      for (uint32_t k = offsets[b]; k < offsets[b + 1]; ++k) {
        Affine<F> p = tab[entries[k] & 0x7fffffffu];
        acc.madd((entries[k] >> 31) && !p.is_inf() ? p.neg() : p);
      }
      return acc.to_affine();
    };

    int rounds = max_rounds;
    if (rounds < 0) { // to completion: ceil(log2(longest bucket)) rounds, plus one that only copies
      rounds = 1;
      while ((1u << rounds) < max_len)
        ++rounds;
      ++rounds;
    }
    int bad = 0;
    if (rounds == 0) return 0; // nothing of the batched path runs

    // buffers exactly as msm_accumulate_batched_enqueue carves them
    const size_t E = (size_t)n_entries, slots0 = ba_slot_bound(E, nb), thr0 = ba_threads_for(slots0);
    std::vector<Affine<F>> pts0(nsel * slots0), pts1(nsel * slots0);
    std::vector<F> prefix(nsel * slots0), totals(nsel * thr0);
    std::vector<uint32_t> off0(nb + 1), off1(nb + 1);
    BaHostExec<F> ex;
    ex.cnt = &cnt;
    BaResult res = ba_run_rounds<F>(
      ex, E, nb, nsel, rounds, entries.data(), tables, offsets.data(), pts0.data(), pts1.data(), prefix.data(), totals.data(),
      off0.data(), off1.data());
    if (ex.overflow) return -2;
    for (int which = 0; which < nsel; ++which) {
      const Affine<F>* cur = (const Affine<F>*)res.cur + (size_t)which * res.pts_stride;
      for (int b = 0; b < nb; ++b) {
        XYZZ<F> acc = XYZZ<F>::inf();
        bool nonempty = ba_bucket_thread(res.off, cur, b, acc);
        if (nonempty != (offsets[b + 1] != offsets[b])) ++bad; // the reduction skips exactly the originally empty buckets
        Affine<F> got = acc.to_affine(), want = reference(tables[which], b);
        if (!(got.x == want.x && got.y == want.y)) ++bad;
      }
    }
    return bad;
  }

} // namespace b200

using namespace b200;

// returns the number of buckets whose batched-affine sum differs from the XYZZ accumulation (0 = agreement),
// -1 for bad arguments, -2 if a buffer or grid bound of the driver would not cover a round; *products_per_add (optional) = field products the model spent per real point addition
extern "C" __attribute__((visibility("default"))) int
b200_batch_affine_selfcheck(int g2, int n_entries, int n_buckets, int max_rounds, unsigned seed, double* products_per_add)
{
  if (n_entries < 0 || n_entries > (1 << 22) || n_buckets < 1 || n_buckets > (1 << 20)) return -1;
  ModelCounts cnt;
  int bad = g2 ? run_batch_affine_model<Fq2>(g2_generator_mont(), n_entries, n_buckets, max_rounds, seed, cnt)
               : run_batch_affine_model<Fq>(g1_generator_mont(), n_entries, n_buckets, max_rounds, seed, cnt);
  if (products_per_add) *products_per_add = cnt.adds ? (double)cnt.products / (double)cnt.adds : 0.0;
  return bad;
}
