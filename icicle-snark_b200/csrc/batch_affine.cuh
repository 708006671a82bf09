// Batched affine addition for bucket accumulation - host/device primitives.
//
// The accumulate kernels run at the field multiplier's peak with 10 products per mixed XYZZ add, so the remaining
// lever is the product count: an affine chord add costs 3 products (lambda * den^-1 is 1, lambda^2 1, y3 1) once the
// inverse of its denominator is known, and Montgomery's trick shares one inversion over a whole batch for 3 more
// products per element (1 running prefix, 2 back-substitution).  Adds inside one bucket are made independent by
// reducing each bucket as a pairwise tree: round r adds elements (2j, 2j+1) of every bucket segment and copies an odd
// tail, so all adds of a round can share inversions.
//
// This header holds the per-pair arithmetic only; csrc/batch_affine_model.cu runs the whole round structure
// (slot -> bucket lookup, per-thread chunks, two-level inversion) on the HOST and checks it against the XYZZ
// accumulation the MSM uses today (tests/test_host_math.py).  No kernel uses it yet: nothing here is on the product path.
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

} // namespace b200
