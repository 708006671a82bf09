// Quad-cooperative group law for the latency-bound top of the bucket reduction (included by msm_impl.cuh).
//
// Above level 0 the reduction is a chain of ~40 dependent point operations on a few thousand threads: with less than one
// warp per SM sub-partition the time of a point operation is the time of its 14 (addition) or 9 (doubling) field products
// issued back to back by ONE lane's warp.  Here the four lanes of a quad hold the same operands; in every step each lane
// multiplies a different pair of them - operands are picked with selects, so the quad executes ONE multiplication
// instruction stream, no divergence - and the four results are exchanged with width-4 shuffles.  An addition is 4 product
// steps instead of 14, a doubling 3 instead of 9; small MSMs (shards of a multi-GPU proof, 100k-constraint circuits), whose
// reduction tail is a third of their run time, get ~2.5x shorter tails.  Same formulas as XYZZ::add / XYZZ::dbl
// (add-2008-s, dbl-2008-s-1) including the exceptional cases, which are quad-uniform because the operands are replicated.
#pragma once
#include "curve.cuh"

namespace b200 {

  template <class Cfg>
  __device__ __forceinline__ Fp<Cfg> quad_shfl(unsigned mask, const Fp<Cfg>& x, int src)
  {
    Fp<Cfg> r;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.v[i] = __shfl_sync(mask, x.v[i], src, 4);
    return r;
  }
  __device__ __forceinline__ Fq2 quad_shfl(unsigned mask, const Fq2& x, int src)
  {
    return {quad_shfl(mask, x.c0, src), quad_shfl(mask, x.c1, src)};
  }
  template <class Cfg>
  __device__ __forceinline__ Fp<Cfg> sel2(bool c, const Fp<Cfg>& a, const Fp<Cfg>& b)
  {
    Fp<Cfg> r;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.v[i] = c ? a.v[i] : b.v[i];
    return r;
  }
  __device__ __forceinline__ Fq2 sel2(bool c, const Fq2& a, const Fq2& b) { return {sel2(c, a.c0, b.c0), sel2(c, a.c1, b.c1)}; }
  // operand of lane `ql` of the quad
  template <class F>
  __device__ __forceinline__ F quad_sel(int ql, const F& a0, const F& a1, const F& a2, const F& a3)
  {
    return sel2((ql & 2) != 0, sel2((ql & 1) != 0, a3, a2), sel2((ql & 1) != 0, a1, a0));
  }
  // one product step: lane k computes xk * yk; out-of-line so the four call sites share one copy of the multiplier
  template <class F>
  __device__ __noinline__ F quad_mul(F x, F y)
  {
    return x * y;
  }

  // a += b by the four lanes of a quad (all hold the same a and b); `ql` = lane within the quad, `qm` = the quad's lane mask
  template <class F>
  __device__ __noinline__ void xyzz_add_quad(XYZZ<F>& a, const XYZZ<F>& b, int ql, unsigned qm)
  {
    if (b.is_inf()) return;
    if (a.is_inf()) {
      a = b;
      return;
    }
    // step 1: u1 = x1 zz2, u2 = x2 zz1, s1 = y1 zzz2, s2 = y2 zzz1
    F t = quad_mul(quad_sel(ql, a.x, b.x, a.y, b.y), quad_sel(ql, b.zz, a.zz, b.zzz, a.zzz));
    const F u1 = quad_shfl(qm, t, 0), u2 = quad_shfl(qm, t, 1), s1 = quad_shfl(qm, t, 2), s2 = quad_shfl(qm, t, 3);
    const F pp_ = u2 - u1, r = s2 - s1;
    if (pp_.is_zero()) {
      if (r.is_zero())
        a = a.dbl_cold(); // P == Q: every lane doubles on its own (rare)
      else
        a = XYZZ<F>::inf();
      return;
    }
    // step 2: pp = p^2, rr = r^2, zzA = zz1 zz2, zzzA = zzz1 zzz2
    t = quad_mul(quad_sel(ql, pp_, r, a.zz, a.zzz), quad_sel(ql, pp_, r, b.zz, b.zzz));
    const F pp = quad_shfl(qm, t, 0), rr = quad_shfl(qm, t, 1), zzA = quad_shfl(qm, t, 2), zzzA = quad_shfl(qm, t, 3);
    // step 3: ppp = p pp, q = u1 pp, zz3 = zzA pp            (lane 3 repeats lane 2's product)
    t = quad_mul(quad_sel(ql, pp_, u1, zzA, zzA), pp);
    const F ppp = quad_shfl(qm, t, 0), q = quad_shfl(qm, t, 1), zz3 = quad_shfl(qm, t, 2);
    const F x3 = rr - ppp - q.dbl();
    // step 4: t1 = r (q - x3), t2 = s1 ppp, zzz3 = zzzA ppp    (lane 3 repeats lane 2's product)
    t = quad_mul(quad_sel(ql, r, s1, zzzA, zzzA), quad_sel(ql, q - x3, ppp, ppp, ppp));
    const F t1 = quad_shfl(qm, t, 0), t2 = quad_shfl(qm, t, 1), zzz3 = quad_shfl(qm, t, 2);
    a.x = x3;
    a.y = t1 - t2;
    a.zz = zz3;
    a.zzz = zzz3;
  }

  // a = 2a by the four lanes of a quad
  template <class F>
  __device__ __noinline__ void xyzz_dbl_quad(XYZZ<F>& a, int ql, unsigned qm)
  {
    if (a.is_inf()) return;
    if (a.y.is_zero()) {
      a = XYZZ<F>::inf();
      return;
    }
    const F u = a.y.dbl();
    // step 1: v = u^2, x2 = x^2                                (lanes 2, 3 repeat)
    F t = quad_mul(quad_sel(ql, u, a.x, u, a.x), quad_sel(ql, u, a.x, u, a.x));
    const F v = quad_shfl(qm, t, 0), x2 = quad_shfl(qm, t, 1);
    const F m = x2.dbl() + x2;
    // step 2: w = u v, s = x v, mm = m^2, zz3 = v zz
    t = quad_mul(quad_sel(ql, u, a.x, m, v), quad_sel(ql, v, v, m, a.zz));
    const F w = quad_shfl(qm, t, 0), s = quad_shfl(qm, t, 1), mm = quad_shfl(qm, t, 2), zz3 = quad_shfl(qm, t, 3);
    const F x3 = mm - s.dbl();
    // step 3: t1 = m (s - x3), t2 = w y, zzz3 = w zzz           (lane 3 repeats lane 2's product)
    t = quad_mul(quad_sel(ql, m, w, w, w), quad_sel(ql, s - x3, a.y, a.zzz, a.zzz));
    const F t1 = quad_shfl(qm, t, 0), t2 = quad_shfl(qm, t, 1), zzz3 = quad_shfl(qm, t, 2);
    a.x = x3;
    a.y = t1 - t2;
    a.zz = zz3;
    a.zzz = zzz3;
  }

} // namespace b200
