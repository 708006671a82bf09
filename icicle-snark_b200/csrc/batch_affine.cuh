// Batched affine addition for bucket accumulation - host/device primitives.
//
// The accumulate kernels run at the field multiplier's peak with 10 products per mixed XYZZ add, so the remaining
// lever is the product count: an affine chord add costs 3 products (lambda * den^-1 is 1, lambda^2 1, y3 1) once the
// inverse of its denominator is known, and Montgomery's trick shares one inversion over a whole batch for 3 more
// products per element (1 running prefix, 2 back-substitution).  Adds inside one bucket are made independent by
// reducing each bucket as a pairwise tree: round r adds elements (2j, 2j+1) of every bucket segment and copies an odd
// tail, so all adds of a round can share inversions.
//
// This header holds the per-pair arithmetic and the per-thread bodies of one round (prefix / invert / finish) as
// host/device functions: csrc/batch_affine_model.cu runs them serially on the HOST and checks the result against the
// XYZZ accumulation the MSM uses today (tests/test_host_math.py); csrc/msm_batch_affine.cuh wraps the same bodies in
// kernels behind B200_BATCH_AFFINE (off by default, not yet run on hardware: not on the product path).
#pragma once
#include "curve.cuh"

namespace b200 {

  enum PairKind : int {
    PAIR_CHORD = 0,   // a != +-b, both finite: lambda = (yb - ya) / (xb - xa)
    PAIR_TANGENT = 1, // a == b: lambda = 3 xa^2 / (2 ya)            (curve coefficient a = 0)
    PAIR_LEFT = 2,    // b is the identity: result a
    PAIR_RIGHT = 3,   // a is the identity: result b
    PAIR_INF = 4      // a == -b (or a 2-torsion point doubled): result is the identity
  };

  // Classifies a + b and returns the denominator whose inverse the add needs (1 for the kinds that need none, so the
  // running product of a batch never becomes zero).
  template <class F>
  B200_HD int pair_prepare(const Affine<F>& a, const Affine<F>& b, F& den)
  {
    den = F::one();
    if (b.is_inf()) return PAIR_LEFT;
    if (a.is_inf()) return PAIR_RIGHT;
    F dx = b.x - a.x;
    if (!dx.is_zero()) {
      den = dx;
      return PAIR_CHORD;
    }
    if (a.y == b.y && !a.y.is_zero()) {
      den = a.y.dbl();
      return PAIR_TANGENT;
    }
    return PAIR_INF;
  }

  // a + b given the inverse of pair_prepare's denominator: 3 products (chord) / 4 + a few additions (tangent)
  template <class F>
  B200_HD Affine<F> pair_finish(int kind, const Affine<F>& a, const Affine<F>& b, const F& den_inv)
  {
    if (kind == PAIR_LEFT) return a;
    if (kind == PAIR_RIGHT) return b;
    if (kind == PAIR_INF) return Affine<F>::inf();
    F num;
    if (kind == PAIR_CHORD) {
      num = b.y - a.y;
    } else {
      F xx = a.x.sqr();
      num = xx.dbl() + xx;
    }
    F lam = num * den_inv;
    F x3 = lam.sqr() - a.x - b.x;
    F y3 = lam * (a.x - x3) - a.y;
    return {x3, y3}; // never (0,0): that point is not on y^2 = x^3 + b, b != 0
  }

  // ------------------------------------------------------------------------------------------------------------
  // One round of the pairwise tree as per-"thread" bodies.  They are plain host/device functions so that the HOST
  // model (batch_affine_model.cu) executes exactly the code the kernels in msm_batch_affine.cuh wrap.
  //
  // A round maps an input of bucket segments `off` (nb + 1 offsets) to an output with segments off_next,
  // len_next = ceil(len / 2): output slot off_next[b] + j = in[off[b] + 2j] + in[off[b] + 2j + 1] (an odd tail is
  // copied).  Thread t owns the BA_M consecutive slots [t BA_M, (t+1) BA_M).
  static constexpr int BA_M = 16;  // slots per thread (first level of Montgomery's trick)
  static constexpr int BA_M2 = 32; // chunk totals per thread of the second level (one true inversion each)

  template <class T>
  B200_HD T ba_ld(const T* p)
  {
#ifdef __CUDA_ARCH__
    constexpr int NQ = sizeof(T) / 16;
    T v;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      d[i] = s[i];
    return v;
#else
    return *p;
#endif
  }
  template <class T>
  B200_HD void ba_st(T* p, const T& v)
  {
#ifdef __CUDA_ARCH__
    constexpr int NQ = sizeof(T) / 16;
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
      d[i] = s[i];
#else
    *p = v;
#endif
  }

  template <class F>
  struct BaRound {
    int round0;               // 1: operands are gathered from the base table through `entries`
    const uint32_t* entries;  // round 0: point index | sign << 31, bucket-sorted
    const Affine<F>* table;   // round 0: base points (Montgomery form)
    const Affine<F>* cur;     // later rounds: the previous round's output
    const uint32_t* off;      // input segments, nb + 1
    const uint32_t* off_next; // output segments; off_next[nb] = number of slots
    int nb;
    F* prefix;                // per slot: product of the thread's denominators in front of it
    F* totals;                // per thread: product of its denominators (pass 1) -> inverse of that product (pass 2)
    Affine<F>* nxt;           // output
  };

  // largest b with off_next[b] <= slot (slot < off_next[nb]): the bucket the slot belongs to
  B200_HD int ba_find_bucket(const uint32_t* off_next, int nb, uint32_t slot)
  {
    int lo = 0, hi = nb; // invariant: off_next[lo] <= slot < off_next[hi]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (off_next[mid] <= slot)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  }

  template <class F>
  B200_HD Affine<F> ba_operand(const BaRound<F>& R, uint32_t k)
  {
    if (!R.round0) return ba_ld(R.cur + k);
    uint32_t e = R.entries[k];
    Affine<F> p = ba_ld(R.table + (e & 0x7fffffffu));
    if (e >> 31) p.y = p.y.neg(); // -(0,0) stays the identity
    return p;
  }

  // the two operands of `slot` in bucket b; returns whether it is a real pair (false: an odd tail, copied)
  template <class F>
  B200_HD bool ba_operands(const BaRound<F>& R, int b, uint32_t slot, Affine<F>& a, Affine<F>& bp)
  {
    uint32_t j = slot - R.off_next[b], L = R.off[b + 1] - R.off[b], src = R.off[b] + 2 * j;
    a = ba_operand(R, src);
    bool pair = 2 * j + 1 < L;
    bp = pair ? ba_operand(R, src + 1) : Affine<F>::inf();
    return pair;
  }

  // pass 1: running products of the denominators of thread t's slots
  template <class F>
  B200_HD void ba_prefix_thread(const BaRound<F>& R, uint32_t t)
  {
    const uint32_t S = R.off_next[R.nb], first = t * BA_M;
    if (first >= S) return;
    const uint32_t last = first + BA_M < S ? first + BA_M : S;
    int b = ba_find_bucket(R.off_next, R.nb, first);
    F run = F::one();
    for (uint32_t s = first; s < last; ++s) {
      while (R.off_next[b + 1] <= s)
        ++b; // next non-empty output segment
      Affine<F> a, bp;
      F den;
      ba_operands(R, b, s, a, bp);
      pair_prepare(a, bp, den);
      ba_st(R.prefix + s, run);
      run = run * den;
    }
    ba_st(R.totals + t, run);
  }

  // pass 2: thread u turns totals[u BA_M2 .. (u+1) BA_M2) into their inverses with one true inversion
  template <class F>
  B200_HD void ba_invert_thread(F* totals, uint32_t n_totals, uint32_t u)
  {
    const uint32_t lo = u * BA_M2;
    if (lo >= n_totals) return;
    const uint32_t hi = lo + BA_M2 < n_totals ? lo + BA_M2 : n_totals;
    F pre[BA_M2];
    F run = F::one();
    for (uint32_t t = lo; t < hi; ++t) {
      pre[t - lo] = run;
      run = run * ba_ld(totals + t);
    }
    F inv = run.inverse();
    for (uint32_t t = hi; t-- > lo;) {
      F tt = ba_ld(totals + t);
      ba_st(totals + t, inv * pre[t - lo]);
      inv = inv * tt;
    }
  }

  // pass 3: thread t walks its slots backwards, recovers each denominator's inverse and finishes the add
  template <class F>
  B200_HD void ba_finish_thread(const BaRound<F>& R, uint32_t t)
  {
    const uint32_t S = R.off_next[R.nb], first = t * BA_M;
    if (first >= S) return;
    const uint32_t last = first + BA_M < S ? first + BA_M : S;
    int b = ba_find_bucket(R.off_next, R.nb, last - 1);
    F inv = ba_ld(R.totals + t);
    for (uint32_t s = last; s-- > first;) {
      while (R.off_next[b] > s)
        --b;
      Affine<F> a, bp;
      F den;
      ba_operands(R, b, s, a, bp);
      int kind = pair_prepare(a, bp, den);
      F den_inv = ba_ld(R.prefix + s) * inv;
      inv = inv * den;
      ba_st(R.nxt + s, pair_finish(kind, a, bp, den_inv));
    }
  }

  // after the last round: bucket b = the (normally single) element left in its segment; longer leftovers
  // (more rounds would have been needed) are summed serially so the result is right for any input
  template <class F>
  B200_HD bool ba_bucket_thread(const uint32_t* off, const Affine<F>* cur, int b, XYZZ<F>& out)
  {
    uint32_t lo = off[b], hi = off[b + 1];
    if (lo == hi) return false; // empty bucket: never read by the reduction
    out = XYZZ<F>::from_affine(ba_ld(cur + lo));
    for (uint32_t k = lo + 1; k < hi; ++k)
      out.madd(ba_ld(cur + k));
    return true;
  }

  // ------------------------------------------------------------------------------------------------------------
  // The round driver, shared by the kernels' host-side launcher (msm_batch_affine.cuh) and the host model: buffer
  // planes for up to BA_MAX_SEL base-point tables that share one sort, ping-pong of the point and offset arrays, and
  // the slot bounds the grids are sized with.  `Exec` runs one pass for `threads` threads of every selection.
  static constexpr int BA_MAX_SEL = 4;

  template <class F>
  struct BaLaunch {
    int round0, nb;
    const uint32_t* entries;
    const Affine<F>* tables[BA_MAX_SEL];
    const Affine<F>* cur;
    Affine<F>* nxt;
    size_t pts_stride; // elements per selection in cur / nxt / prefix
    const uint32_t* off;
    const uint32_t* off_next;
    F* prefix;
    F* totals;
    size_t tot_stride; // elements per selection in totals
  };

  template <class F>
  B200_HD BaRound<F> ba_round_of(const BaLaunch<F>& L, int which)
  {
    BaRound<F> R;
    R.round0 = L.round0;
    R.entries = L.entries;
    R.table = L.tables[which];
    R.cur = L.cur + (size_t)which * L.pts_stride;
    R.off = L.off;
    R.off_next = L.off_next;
    R.nb = L.nb;
    R.prefix = L.prefix + (size_t)which * L.pts_stride;
    R.totals = L.totals + (size_t)which * L.tot_stride;
    R.nxt = L.nxt + (size_t)which * L.pts_stride;
    return R;
  }

  // a round's output has at most ceil(E_in / 2) + nb slots (one copied tail per bucket)
  inline size_t ba_slot_bound(size_t e_in, int nb) { return (e_in + 1) / 2 + (size_t)nb; }
  inline size_t ba_threads_for(size_t slots) { return (slots + BA_M - 1) / BA_M; }

  struct BaResult { // where the last round left the bucket segments
    const uint32_t* off;
    const void* cur;
    size_t pts_stride;
  };

  // pts0/pts1: nsel * ba_slot_bound(E, nb) points each; prefix: as many field elements; totals: nsel *
  // ba_threads_for(that bound); off0/off1: nb + 1 words each
  template <class F, class Exec>
  BaResult ba_run_rounds(
    Exec& ex, size_t E, int nb, int nsel, int rounds, const uint32_t* entries, const Affine<F>* const* tables,
    const uint32_t* offsets0, Affine<F>* pts0, Affine<F>* pts1, F* prefix, F* totals, uint32_t* off0, uint32_t* off1)
  {
    const size_t slots0 = ba_slot_bound(E, nb);
    Affine<F>* pts[2] = {pts0, pts1};
    uint32_t* offs[2] = {off0, off1};
    BaLaunch<F> L;
    L.nb = nb;
    L.entries = entries;
    for (int k = 0; k < BA_MAX_SEL; ++k)
      L.tables[k] = tables[k < nsel ? k : 0];
    L.pts_stride = slots0;
    L.prefix = prefix;
    L.totals = totals;
    L.tot_stride = ba_threads_for(slots0);
    const uint32_t* off = offsets0;
    const Affine<F>* cur = pts[1]; // not read in round 0
    size_t in_bound = E;
    for (int r = 0; r < rounds; ++r) {
      uint32_t* off_next = offs[r & 1];
      Affine<F>* nxt = pts[r & 1];
      ex.next_offsets(off, nb, off_next); // off_next = exclusive scan of ceil(len / 2), total in off_next[nb]
      const size_t out_bound = ba_slot_bound(in_bound, nb);
      const size_t threads = ba_threads_for(out_bound);
      L.round0 = r == 0;
      L.cur = cur;
      L.nxt = nxt;
      L.off = off;
      L.off_next = off_next;
      ex.prefix(L, threads, nsel);
      ex.invert(L, (threads + BA_M2 - 1) / BA_M2, nsel);
      ex.finish(L, threads, nsel);
      off = off_next;
      cur = nxt;
      if (out_bound < in_bound) in_bound = out_bound;
    }
    return {off, cur, slots0};
  }

} // namespace b200
