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

  template <class F>
  static int run_batch_affine_model(const Affine<F>& gen, int n_entries, int nb, int max_rounds, uint32_t seed, ModelCounts& cnt)
  {
    int rounds_done = 0;
    std::mt19937 rng(seed);
    // base points: small multiples of the generator, so equal points, opposite points and the identity all occur
    const int n_pts = 24;
    std::vector<Affine<F>> table(n_pts);
    {
      XYZZ<F> g = XYZZ<F>::from_affine(gen), acc = XYZZ<F>::inf();
      for (int i = 0; i < n_pts; ++i) {
        if (i % 8 != 7) acc.add(g);                       // every 8th point repeats its predecessor
        table[i] = (i % 11 == 10) ? Affine<F>::inf() : acc.to_affine(); // a few identities (zkeys do contain (0,0))
      }
    }
    // entries (point index | sign << 31), bucket-sorted, with offsets - the MSM's sort output
    std::vector<uint32_t> bucket_of(n_entries), offsets(nb + 1, 0);
    for (int k = 0; k < n_entries; ++k) {
      // skewed: a quarter of the entries land in bucket 0, some buckets stay empty
      uint32_t b = (rng() & 3) == 0 ? 0 : rng() % (uint32_t)nb;
      if (nb > 4 && b == 3) b = 2;
      bucket_of[k] = b;
      ++offsets[b + 1];
    }
    for (int b = 0; b < nb; ++b)
      offsets[b + 1] += offsets[b];
    std::vector<uint32_t> entries(n_entries), cursor(offsets.begin(), offsets.end() - 1);
    for (int k = 0; k < n_entries; ++k)
      entries[cursor[bucket_of[k]]++] = (rng() % n_pts) | ((rng() & 1u) << 31);

    // reference: what msm_accumulate_kernel computes (XYZZ += affine)
    std::vector<Affine<F>> want(nb);
    for (int b = 0; b < nb; ++b) {
      XYZZ<F> acc = XYZZ<F>::inf();
      for (uint32_t k = offsets[b]; k < offsets[b + 1]; ++k) {
        Affine<F> p = table[entries[k] & 0x7fffffffu];
        acc.madd((entries[k] >> 31) && !p.is_inf() ? p.neg() : p);
      }
      want[b] = acc.to_affine();
    }

    // ---- the rounds: the per-thread bodies of batch_affine.cuh (the ones the kernels wrap), one "thread" after another
    std::vector<Affine<F>> cur, nxt;
    std::vector<uint32_t> off = offsets, off_next(nb + 1);
    std::vector<F> prefix, totals;
    bool first = true;
    for (;;) {
      uint32_t max_len = 0;
      off_next[0] = 0;
      for (int b = 0; b < nb; ++b) {
        uint32_t L = off[b + 1] - off[b];
        max_len = std::max(max_len, L);
        off_next[b + 1] = off_next[b] + (L + 1) / 2;
      }
      if (max_len <= 1 && !first) break;
      if (max_rounds >= 0 && rounds_done >= max_rounds) break; // leftovers are summed by ba_bucket_thread
      const uint32_t S = off_next[nb], T = (S + BA_M - 1) / BA_M, U = (T + BA_M2 - 1) / BA_M2;
      nxt.assign(S, Affine<F>::inf());
      prefix.assign(S ? S : 1, F::zero());
      totals.assign(T ? T : 1, F::zero());
      BaRound<F> R;
      R.round0 = first ? 1 : 0;
      R.entries = entries.data();
      R.table = table.data();
      R.cur = cur.data();
      R.off = off.data();
      R.off_next = off_next.data();
      R.nb = nb;
      R.prefix = prefix.data();
      R.totals = totals.data();
      R.nxt = nxt.data();
      for (uint32_t t = 0; t < T + 2; ++t) // two threads past the end, as a rounded-up grid has
        ba_prefix_thread(R, t);
      for (uint32_t u = 0; u < U + 2; ++u)
        ba_invert_thread(totals.data(), T, u);
      for (uint32_t t = T + 2; t-- > 0;)
        ba_finish_thread(R, t);
      // accounting (not part of the kernels): real adds of this round and the products they cost
      for (uint32_t sl = 0; sl < S; ++sl) {
        int b = ba_find_bucket(off_next.data(), nb, sl);
        Affine<F> a, bp;
        F den;
        bool pair = ba_operands(R, b, sl, a, bp);
        int kind = pair_prepare(a, bp, den);
        if (pair && kind <= PAIR_TANGENT) {
          ++cnt.adds;
          cnt.products += kind == PAIR_CHORD ? 3 : 4;
        }
      }
      cnt.products += 3 * (uint64_t)S + 3 * (uint64_t)T + 384 * (uint64_t)U;
      cnt.inversions += U;
      cur.swap(nxt);
      off = off_next;
      first = false;
      ++rounds_done;
    }

    int bad = 0;
    for (int b = 0; b < nb; ++b) {
      XYZZ<F> acc = XYZZ<F>::inf();
      if (first) { // no round ran (max_rounds == 0): the buckets come straight from the gathered entries
        for (uint32_t k = offsets[b]; k < offsets[b + 1]; ++k) {
          Affine<F> p = table[entries[k] & 0x7fffffffu];
          acc.madd((entries[k] >> 31) && !p.is_inf() ? p.neg() : p);
        }
      } else {
        ba_bucket_thread(off.data(), cur.data(), b, acc);
      }
      Affine<F> got = acc.to_affine();
      if (!(got.x == want[b].x && got.y == want[b].y)) ++bad;
    }
    return bad;
  }

} // namespace b200

using namespace b200;

// returns the number of buckets whose batched-affine sum differs from the XYZZ accumulation (0 = agreement),
// -1 for bad arguments; *products_per_add (optional) = field products the model spent per real point addition
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
