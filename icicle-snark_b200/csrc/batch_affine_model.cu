// HOST model of bucket accumulation by batched affine addition - the round structure the
// round-2 kernels will have, executed serially and checked against the XYZZ accumulation the MSM uses today.
//
//   round r:  every bucket segment of L elements becomes ceil(L/2): slot j <- in[2j] + in[2j+1], an odd tail is copied
//   pass 1:   "thread" t owns slots [t m, (t+1) m): classifies each pair, stores the running product of the
//             denominators in front of it (32 B per slot) and its chunk total
//   pass 2:   the chunk totals are inverted with a second level of the same trick (one true inversion per m2 chunks)
//   pass 3:   thread t walks its slots backwards: inverse of its denominator = prefix * running inverse, finishes the add
//
// Field products per add: 1 (prefix) + 2 (back-substitution) + 3 (chord) + (3 + I/m2)/m for the second level, I = one
// inversion (~310 products): 6.3 at m = 16, m2 = 64.  Today's mixed XYZZ add costs 10.
// Test/model code: exported as b200_batch_affine_selfcheck for tests/test_host_math.py; no kernel calls it.
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
  static int run_batch_affine_model(const Affine<F>& gen, int n_entries, int nb, int m, int m2, uint32_t seed, ModelCounts& cnt)
  {
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

    // ---- the rounds
    std::vector<Affine<F>> cur, nxt;
    std::vector<uint32_t> off = offsets, off_next(nb + 1);
    bool first = true;
    auto input = [&](uint32_t k) -> Affine<F> {
      if (!first) return cur[k];
      Affine<F> p = table[entries[k] & 0x7fffffffu]; // round 0 gathers through the sorted entries
      return (entries[k] >> 31) && !p.is_inf() ? p.neg() : p;
    };
    for (;;) {
      uint32_t max_len = 0;
      off_next[0] = 0;
      for (int b = 0; b < nb; ++b) {
        uint32_t L = off[b + 1] - off[b];
        max_len = std::max(max_len, L);
        off_next[b + 1] = off_next[b] + (L + 1) / 2;
      }
      if (max_len <= 1 && !first) break;
      const uint32_t S = off_next[nb];
      nxt.assign(S, Affine<F>::inf());
      // slot -> (bucket, pair) exactly as a kernel would: binary search in the next round's offsets
      auto operands = [&](uint32_t slot, Affine<F>& a, Affine<F>& bpt) {
        int b = (int)(std::upper_bound(off_next.begin(), off_next.end(), slot) - off_next.begin()) - 1;
        uint32_t j = slot - off_next[b], L = off[b + 1] - off[b], src = off[b] + 2 * j;
        a = input(src);
        bpt = (2 * j + 1 < L) ? input(src + 1) : Affine<F>::inf(); // odd tail: copied
        return 2 * j + 1 < L;
      };
      const uint32_t T = (S + m - 1) / m;
      std::vector<F> prefix(S), totals(T), inv_totals(T);
      for (uint32_t t = 0; t < T; ++t) { // pass 1
        F run = F::one();
        for (uint32_t s = t * m; s < std::min<uint32_t>(S, (t + 1) * m); ++s) {
          Affine<F> a, bpt;
          F den;
          bool real = operands(s, a, bpt);
          int kind = pair_prepare(a, bpt, den);
          prefix[s] = run;
          run = run * den;
          ++cnt.products;
          if (real && kind <= PAIR_TANGENT) ++cnt.adds;
        }
        totals[t] = run;
      }
      for (uint32_t u = 0; u * m2 < T; ++u) { // pass 2: second level of the trick over the chunk totals
        uint32_t lo = u * m2, hi = std::min<uint32_t>(T, lo + m2);
        std::vector<F> pre(hi - lo);
        F run = F::one();
        for (uint32_t t = lo; t < hi; ++t) {
          pre[t - lo] = run;
          run = run * totals[t];
        }
        F inv = run.inverse();
        ++cnt.inversions;
        for (uint32_t t = hi; t-- > lo;) {
          inv_totals[t] = inv * pre[t - lo];
          inv = inv * totals[t];
        }
        cnt.products += 3 * (uint64_t)(hi - lo);
      }
      for (uint32_t t = 0; t < T; ++t) { // pass 3
        F inv = inv_totals[t];
        for (uint32_t s = std::min<uint32_t>(S, (t + 1) * m); s-- > t * m;) {
          Affine<F> a, bpt;
          F den;
          operands(s, a, bpt);
          int kind = pair_prepare(a, bpt, den);
          F den_inv = prefix[s] * inv;
          inv = inv * den;
          cnt.products += 2;
          if (kind <= PAIR_TANGENT) cnt.products += 3;
          nxt[s] = pair_finish(kind, a, bpt, den_inv);
        }
      }
      cur.swap(nxt);
      off = off_next;
      first = false;
    }

    int bad = 0;
    for (int b = 0; b < nb; ++b) {
      uint32_t L = off[b + 1] - off[b];
      Affine<F> got = L ? cur[off[b]] : Affine<F>::inf();
      if (!(got.x == want[b].x && got.y == want[b].y)) ++bad;
    }
    return bad;
  }

} // namespace b200

using namespace b200;

// returns the number of buckets whose batched-affine sum differs from the XYZZ accumulation (0 = agreement),
// -1 for bad arguments; *products_per_add (optional) = field products the model spent per real point addition
extern "C" __attribute__((visibility("default"))) int
b200_batch_affine_selfcheck(int g2, int n_entries, int n_buckets, int m, int m2, unsigned seed, double* products_per_add)
{
  if (n_entries < 0 || n_entries > (1 << 22) || n_buckets < 1 || n_buckets > (1 << 20) || m < 1 || m2 < 1) return -1;
  ModelCounts cnt;
  int bad = g2 ? run_batch_affine_model<Fq2>(g2_generator_mont(), n_entries, n_buckets, m, m2, seed, cnt)
               : run_batch_affine_model<Fq>(g1_generator_mont(), n_entries, n_buckets, m, m2, seed, cnt);
  if (products_per_add) *products_per_add = cnt.adds ? (double)cnt.products / (double)cnt.adds : 0.0;
  return bad;
}
