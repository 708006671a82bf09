// BN254 G1/G2 group arithmetic for the MSM path.
//
// The reference accumulates in homogeneous projective coordinates with the complete RCB formulas
//   /root/reference/icicle/include/icicle/curves/projective.h:82-124 (add), 128-169 (mixed), 54-80 (dbl).
// Here buckets are XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 == ZZZ^2; ZZ == 0 is the identity) and base points
// are affine: mixed add = 8M+2S, full add = 12M+2S, with the P==Q / P==-Q / identity cases made
// explicit (the RCB formulas never see them; real zkeys do contain repeated and (0,0) points).
// At the API boundary results are converted to the reference's layout:
// homogeneous (X:Y:Z), identity (0,1,0), standard-form limbs  (projective.h:26-38).
#pragma once
#include "field.cuh"

namespace b200 {

  template <class F>
  struct alignas(16) Affine {
    F x, y;
    // (0,0) encodes the point at infinity (affine.h)
    B200_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static B200_HD Affine inf() { return {F::zero(), F::zero()}; }
    B200_HD Affine neg() const { return {x, y.neg()}; }
  };

  template <class F>
  struct alignas(16) Projective { // reference boundary layout
    F x, y, z;
  };

  template <class F>
  struct alignas(16) XYZZ {
    F x, y, zz, zzz;

    static B200_HD XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    B200_HD bool is_inf() const { return zz.is_zero(); }
    static B200_HD XYZZ from_affine(const Affine<F>& p)
    {
      if (p.is_inf()) return inf();
      return {p.x, p.y, F::one(), F::one()};
    }
    B200_HD XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

    // out-of-line twins for the exceptional branches of madd/add (P == Q is rare): keeps the hot
    // accumulate loop's register and instruction footprint to the generic case
#if defined(__CUDACC__)
    static __host__ __device__ __noinline__ XYZZ dbl_affine_cold(const Affine<F>& p) { return dbl_affine(p); }
    __host__ __device__ __noinline__ XYZZ dbl_cold() const { return dbl(); }
#else
    static XYZZ dbl_affine_cold(const Affine<F>& p) { return dbl_affine(p); }
    XYZZ dbl_cold() const { return dbl(); }
#endif

    // 2*(affine p), p != inf            (mdbl-2008-s-1)
    static B200_HD XYZZ dbl_affine(const Affine<F>& p)
    {
      if (p.y.is_zero()) return inf(); // order-2 point: none on BN254, kept for totality
      F u = p.y.dbl();
      F v = u.sqr();
      F w = u * v;
      F s = p.x * v;
      F x2 = p.x.sqr();
      F m = x2.dbl() + x2;
      F x3 = m.sqr() - s.dbl();
      F y3 = m * (s - x3) - w * p.y;
      return {x3, y3, v, w};
    }

    // 2*this                             (dbl-2008-s-1)
    B200_HD XYZZ dbl() const
    {
      if (is_inf() || y.is_zero()) return inf();
      F u = y.dbl();
      F v = u.sqr();
      F w = u * v;
      F s = x * v;
      F x2 = x.sqr();
      F m = x2.dbl() + x2;
      F x3 = m.sqr() - s.dbl();
      F y3 = m * (s - x3) - w * y;
      return {x3, y3, v * zz, w * zzz};
    }

    // this += affine p                   (madd-2008-s, with the exceptional cases)
    B200_HD void madd(const Affine<F>& p)
    {
      if (p.is_inf()) return;
      if (is_inf()) {
        x = p.x;
        y = p.y;
        zz = F::one();
        zzz = F::one();
        return;
      }
      F u2 = p.x * zz;
      F s2 = p.y * zzz;
      F pp_ = u2 - x;
      F r = s2 - y;
      if (pp_.is_zero()) {
        if (r.is_zero())
          *this = dbl_affine_cold(p);
        else
          *this = inf();
        return;
      }
      F pp = pp_.sqr();
      F ppp = pp_ * pp;
      F q = x * pp;
      F x3 = r.sqr() - ppp - q.dbl();
      y = F::mul_sub(r, q - x3, y, ppp);
      x = x3;
      zz = zz * pp;
      zzz = zzz * ppp;
    }

    // this += o                          (add-2008-s, with the exceptional cases)
    B200_HD void add(const XYZZ& o)
    {
      if (o.is_inf()) return;
      if (is_inf()) {
        *this = o;
        return;
      }
      F u1 = x * o.zz;
      F u2 = o.x * zz;
      F s1 = y * o.zzz;
      F s2 = o.y * zzz;
      F pp_ = u2 - u1;
      F r = s2 - s1;
      if (pp_.is_zero()) {
        if (r.is_zero())
          *this = dbl_cold();
        else
          *this = inf();
        return;
      }
      F pp = pp_.sqr();
      F ppp = pp_ * pp;
      F q = u1 * pp;
      F x3 = r.sqr() - ppp - q.dbl();
      y = F::mul_sub(r, q - x3, s1, ppp);
      x = x3;
      zz = zz * o.zz * pp;
      zzz = zzz * o.zzz * ppp;
    }

    // -> homogeneous projective, still Montgomery. x = X/ZZ = X*ZZZ/(ZZ*ZZZ), y = Y/ZZZ = Y*ZZ/(ZZ*ZZZ)
    B200_HD Projective<F> to_projective() const
    {
      if (is_inf()) return {F::zero(), F::one(), F::zero()};
      return {x * zzz, y * zz, zz * zzz};
    }

    B200_HD Affine<F> to_affine() const
    {
      if (is_inf()) return Affine<F>::inf();
      // 1/ZZ = ZZ^2/ZZZ^2 * ... : simply invert both denominators with one inversion
      F d = (zz * zzz).inverse(); // 1/(ZZ*ZZZ)
      return {x * zzz * d, y * zz * d};
    }
  };

  template <class F>
  B200_HD XYZZ<F> xyzz_from_projective(const Projective<F>& p)
  {
    // (X:Y:Z) homogeneous -> XYZZ with ZZ = Z^2, ZZZ = Z^3: x = X/Z = XZ/Z^2, y = Y/Z = YZ^2/Z^3
    if (p.z.is_zero()) return XYZZ<F>::inf();
    F z2 = p.z.sqr();
    return {p.x * p.z, p.y * z2, z2, z2 * p.z};
  }

  typedef Affine<Fq> G1Affine;
  typedef Affine<Fq2> G2Affine;
  typedef XYZZ<Fq> G1XYZZ;
  typedef XYZZ<Fq2> G2XYZZ;
  typedef Projective<Fq> G1Projective;
  typedef Projective<Fq2> G2Projective;

} // namespace b200
