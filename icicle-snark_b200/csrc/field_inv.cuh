// Modular inversion by Bernstein-Yang division steps ("safegcd", TCHES 2019) on 30-bit signed limbs, for the batched
// affine bucket accumulation (msm_batch_affine.cuh), where one inversion is shared by a few additions per thread.
//
// Why not Fp::inverse(): Fermat's a^(p-2) is ~380 Montgomery products on the integer multiply pipe - the pipe the MSM
// is bound by.  Here 600 division steps run on the low words in the ALU pipe (shifts, adds, selects; the pipe the
// field multiplier leaves mostly idle) in 20 batches of 30; each batch is folded into a 2x2 integer matrix applied to
// the full-width (f, g) and to the Bezout cofactors (d, e) with 90 wide multiply-adds: 1800 in total, the cost of ~13
// Montgomery products.  Straight-line, no data-dependent branches: all lanes of a warp stay converged.
//
// Invariants (mod p):  d * x = f * K,  e * x = g * K  with K = R^2, so that for a Montgomery-form input xR the result
// d = (xR)^-1 R^2 = x^-1 R comes out in Montgomery form directly.  590 half-delta division steps suffice for any
// 256-bit modulus (Bernstein-Yang bound as computed for the 256-bit case by the libsecp256k1 authors); 600 are run.
// inverse(0) = 0, like Fp::inverse() and the reference (modular_arithmetic.h:603).
#pragma once
#include "field.cuh"

namespace b200 {

  namespace safegcd {
    static constexpr int32_t M30 = (1 << 30) - 1;

    struct Limbs9 {
      int32_t v[9];
    };

    // 8x32-bit little-endian words (< 2^256) -> 9 limbs of 30 bits
    template <class Get>
    constexpr B200_HD Limbs9 to30(Get w)
    {
      Limbs9 r{};
      for (int i = 0; i < 9; ++i) {
        const int bit = 30 * i, wd = bit >> 5, sh = bit & 31;
        uint64_t x = wd < 8 ? (uint64_t)w(wd) : 0;
        if (wd + 1 < 8) x |= (uint64_t)w(wd + 1) << 32;
        r.v[i] = (int32_t)((x >> sh) & (uint32_t)M30);
      }
      return r;
    }

    template <class Cfg>
    struct Consts {
      static constexpr B200_HD Limbs9 p30()
      {
        return to30([](int i) { return Cfg::P(i); });
      }
      static constexpr B200_HD Limbs9 r2_30()
      {
        return to30([](int i) { return Cfg::R2(i); });
      }
      // p^-1 mod 2^30 (Newton: five doublings of precision from 1, p odd)
      static constexpr B200_HD uint32_t pinv30()
      {
        uint32_t p0 = Cfg::P(0), x = 1;
        for (int i = 0; i < 5; ++i)
          x *= 2u - p0 * x;
        return x & (uint32_t)M30;
      }
    };

    // 30 division steps on the low words; returns the transition matrix scaled by 2^30:
    // 2^30 (f', g') = [[u, v], [q, r]] (f, g).  D = 2*delta (odd).
    B200_HD void divsteps30(int32_t& D, uint32_t f, uint32_t g, int32_t& u, int32_t& v, int32_t& q, int32_t& r)
    {
      u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
      for (int i = 0; i < 30; ++i) {
        const bool odd = g & 1u;
        const bool swap = odd & (D > 0);
        // operand added to the second row / to g: +row1, -row1 (swap) or nothing (g even)
        const uint32_t fa = swap ? 0u - f : f;
        const int32_t ua = swap ? -u : u, va = swap ? -v : v;
        const uint32_t gs = g + (odd ? fa : 0u);
        const int32_t qn = q + (odd ? ua : 0), rn = r + (odd ? va : 0);
        f = swap ? g : f;
        u = (swap ? q : u) << 1;
        v = (swap ? r : v) << 1;
        g = gs >> 1;
        q = qn;
        r = rn;
        D = (swap ? -D : D) + 2;
      }
    }

    // (f, g) <- [[u, v], [q, r]] (f, g) / 2^30 (exact)
    B200_HD void update_fg(Limbs9& f, Limbs9& g, int32_t u, int32_t v, int32_t q, int32_t r)
    {
      int64_t cf = (int64_t)u * f.v[0] + (int64_t)v * g.v[0];
      int64_t cg = (int64_t)q * f.v[0] + (int64_t)r * g.v[0];
      cf >>= 30;
      cg >>= 30;
#pragma unroll
      for (int i = 1; i < 9; ++i) {
        cf += (int64_t)u * f.v[i] + (int64_t)v * g.v[i];
        cg += (int64_t)q * f.v[i] + (int64_t)r * g.v[i];
        f.v[i - 1] = (int32_t)cf & M30;
        g.v[i - 1] = (int32_t)cg & M30;
        cf >>= 30;
        cg >>= 30;
      }
      f.v[8] = (int32_t)cf;
      g.v[8] = (int32_t)cg;
    }

    // (d, e) <- [[u, v], [q, r]] (d, e) / 2^30 mod p, kept in (-2p, p)
    template <class Cfg>
    B200_HD void update_de(Limbs9& d, Limbs9& e, int32_t u, int32_t v, int32_t q, int32_t r)
    {
      constexpr Limbs9 P = Consts<Cfg>::p30();
      constexpr uint32_t PINV = Consts<Cfg>::pinv30();
      const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
      int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
      int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
      int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
      // multiples of p that clear the low 30 bits
      md -= (int32_t)((PINV * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      me -= (int32_t)((PINV * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
      cd += (int64_t)P.v[0] * md;
      ce += (int64_t)P.v[0] * me;
      cd >>= 30;
      ce >>= 30;
#pragma unroll
      for (int i = 1; i < 9; ++i) {
        cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i] + (int64_t)P.v[i] * md;
        ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i] + (int64_t)P.v[i] * me;
        d.v[i - 1] = (int32_t)cd & M30;
        e.v[i - 1] = (int32_t)ce & M30;
        cd >>= 30;
        ce >>= 30;
      }
      d.v[8] = (int32_t)cd;
      e.v[8] = (int32_t)ce;
    }

    // d in (-2p, p) -> sign * d mod p in [0, p), as 8x32-bit words
    template <class Cfg>
    B200_HD void finish(Limbs9 d, int32_t negate_mask, uint32_t (&out)[8])
    {
      constexpr Limbs9 P = Consts<Cfg>::p30();
      // + p if negative; then conditional negation; then + p again if negative
      int32_t add = d.v[8] >> 31;
#pragma unroll
      for (int i = 0; i < 9; ++i)
        d.v[i] = ((d.v[i] + (P.v[i] & add)) ^ negate_mask) - negate_mask;
      // carry propagation (limbs may be out of [0, 2^30) now)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d.v[i + 1] += d.v[i] >> 30;
        d.v[i] &= M30;
      }
      add = d.v[8] >> 31;
#pragma unroll
      for (int i = 0; i < 9; ++i)
        d.v[i] += P.v[i] & add;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d.v[i + 1] += d.v[i] >> 30;
        d.v[i] &= M30;
      }
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int bit = 32 * w, l = bit / 30, sh = bit % 30; // word w = bits [32w, 32w+32)
        uint64_t x = (uint64_t)(uint32_t)d.v[l] >> sh;
        x |= (uint64_t)(uint32_t)d.v[l + 1] << (30 - sh);
        if (60 - sh < 32 && l + 2 < 9) x |= (uint64_t)(uint32_t)d.v[l + 2] << (60 - sh);
        out[w] = (uint32_t)x;
      }
    }
  } // namespace safegcd

  // x^-1 for a Montgomery-form x (result in Montgomery form); 0 -> 0
  template <class Cfg>
  B200_HD Fp<Cfg> inverse_safegcd(const Fp<Cfg>& x)
  {
    using namespace safegcd;
    Limbs9 f = Consts<Cfg>::p30();
    Limbs9 e = Consts<Cfg>::r2_30();
    Limbs9 g = to30([&x](int i) { return x.v[i]; });
    Limbs9 d{};
    int32_t D = 1;
#pragma unroll 1
    for (int it = 0; it < 20; ++it) {
      int32_t u, v, q, r;
      const uint32_t f0 = (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30);
      const uint32_t g0 = (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30);
      divsteps30(D, f0, g0, u, v, q, r);
      update_de<Cfg>(d, e, u, v, q, r);
      update_fg(f, g, u, v, q, r);
    }
    // g = 0 now and f = +-gcd = +-1 (f = +-p when x = 0, where d = 0 anyway)
    Fp<Cfg> out;
    finish<Cfg>(d, f.v[8] >> 31, out.v);
    return out;
  }

} // namespace b200
