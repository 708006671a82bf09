// Batched affine addition for bucket accumulation - the per-pair arithmetic (host/device).
//
// The XYZZ accumulate kernel runs at the integer multiply pipe's peak with 10 products per mixed add, so the remaining
// lever is the product count: an affine chord add costs 3 products (lambda = num * den^-1, lambda^2, y3) once the
// inverse of its denominator is known, and Montgomery's trick shares one inversion over a batch for 3 more products per
// element (1 running prefix, 2 back-substitution).  Adds inside one bucket are made independent by reducing each bucket
// as a pairwise tree: a round adds elements (2j, 2j+1) of every bucket segment and copies an odd tail.  The kernels
// and the round driver are in msm_batch_affine.cuh; the shared inversion is field_inv.cuh.
#pragma once
#include "curve.cuh"
#include "field_inv.cuh"

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

  // inversion shared by a batch: division steps for the base field, norm + one base-field inversion for Fq2
  B200_HD Fq batch_inverse(const Fq& x) { return inverse_safegcd(x); }
  B200_HD Fq2 batch_inverse(const Fq2& x)
  {
    Fq t = inverse_safegcd(x.c0.sqr() + x.c1.sqr());
    return {x.c0 * t, (x.c1 * t).neg()};
  }

} // namespace b200
