// BN254 prime-field arithmetic (Fr and Fq) on 8x32-bit limbs, Montgomery form, R = 2^256.
//
// Replaces the reference's Field<CONFIG> (Karatsuba + Barrett):
//   /root/reference/icicle/include/icicle/math/modular_arithmetic.h:337-352,383-388,490-494,556-564,601-631
//   /root/reference/icicle/backend/cuda/include/cuda_math.h:299-346,491-526
// Values held in Fp<> are ALWAYS Montgomery residues, fully reduced to [0, p).  The reference's
// boundary format (standard form, 8xu32 LE) is produced/consumed by to_mont()/from_mont().
//
// Device path: generated PTX carry chains (field_asm.inc.h, see tools/gen_field_asm.py).
// Host path (used by the host-side helpers bn254_add/ecadd/..., never as a fallback for
// device work): portable 64-bit CIOS.
#pragma once
#include <cstdint>
#include <cstring>

#include "field_asm.inc.h"

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__
#elif defined(__CUDACC__)
// host pass of nvcc: let g++ decide (force-inlining the whole tower into the epilogue costs minutes of build)
#define B200_HD __host__ __device__ inline
#define B200_D __device__ inline
#else
#define B200_HD inline
#define B200_D inline
#endif

namespace b200 {

  struct FrCfg {
    // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  (bn254_scalar.h:9-10)
    static constexpr B200_HD uint32_t P(int i)
    {
      constexpr uint32_t t[8] = {0xf0000001, 0x43e1f593, 0x79b97091, 0x2833e848,
                                      0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};
      return t[i];
    }
    static constexpr B200_HD uint32_t ONE(int i)
    {
      constexpr uint32_t t[8] = {0x4ffffffb, 0xac96341c, 0x9f60cd29, 0x36fc7695,
                                        0x7879462e, 0x666ea36f, 0x9a07df2f, 0x0e0a77c1};
      return t[i];
    } // 2^256 mod r
    static constexpr B200_HD uint32_t R2(int i)
    {
      constexpr uint32_t t[8] = {0xae216da7, 0x1bb8e645, 0xe35c59e3, 0x53fe3ab1,
                                       0x53bb8085, 0x8c49833d, 0x7f4e44a5, 0x0216d0b1};
      return t[i];
    } // 2^512 mod r
    static constexpr uint32_t INV = 0xefffffff; // -r^-1 mod 2^32
#ifdef __CUDA_ARCH__
    static B200_D void redc(uint32_t (&X)[8], uint32_t (&Y)[8]) { ptx::redc_fr(X, Y); }
    static B200_D uint32_t subp(uint32_t (&T)[8], const uint32_t (&R)[8]) { return ptx::subp_fr(T, R); }
    static B200_D void addp_masked(uint32_t (&R)[8], uint32_t mk) { ptx::addp_masked_fr(R, mk); }
    // Montgomery reduction of a 16-limb value held as X + Y*2^32 (< p*2^256): eight SOS steps
    static B200_D void redc16(uint32_t (&X)[16], uint32_t (&Y)[16], uint32_t (&K)[8], uint32_t& c)
    {
      ptx::redc16_step0_fr(X, Y, K, c);
      ptx::redc16_step1_fr(X, Y, K, c);
      ptx::redc16_step2_fr(X, Y, K, c);
      ptx::redc16_step3_fr(X, Y, K, c);
      ptx::redc16_step4_fr(X, Y, K, c);
      ptx::redc16_step5_fr(X, Y, K, c);
      ptx::redc16_step6_fr(X, Y, K, c);
      ptx::redc16_step7_fr(X, Y, K, c);
    }
    static B200_D void addp2(uint32_t (&T)[16]) { ptx::addp2_fr(T); }
#endif
  };

  struct FqCfg {
    // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  (bn254_base.h:8-9)
    static constexpr B200_HD uint32_t P(int i)
    {
      constexpr uint32_t t[8] = {0xd87cfd47, 0x3c208c16, 0x6871ca8d, 0x97816a91,
                                      0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};
      return t[i];
    }
    static constexpr B200_HD uint32_t ONE(int i)
    {
      constexpr uint32_t t[8] = {0xc58f0d9d, 0xd35d438d, 0xf5c70b3d, 0x0a78eb28,
                                        0x7879462c, 0x666ea36f, 0x9a07df2f, 0x0e0a77c1};
      return t[i];
    } // 2^256 mod q
    static constexpr B200_HD uint32_t R2(int i)
    {
      constexpr uint32_t t[8] = {0x538afa89, 0xf32cfc5b, 0xd44501fb, 0xb5e71911,
                                       0x0a417ff6, 0x47ab1eff, 0xcab8351f, 0x06d89f71};
      return t[i];
    } // 2^512 mod q
    static constexpr uint32_t INV = 0xe4866389; // -q^-1 mod 2^32
#ifdef __CUDA_ARCH__
    static B200_D void redc(uint32_t (&X)[8], uint32_t (&Y)[8]) { ptx::redc_fq(X, Y); }
    static B200_D uint32_t subp(uint32_t (&T)[8], const uint32_t (&R)[8]) { return ptx::subp_fq(T, R); }
    static B200_D void addp_masked(uint32_t (&R)[8], uint32_t mk) { ptx::addp_masked_fq(R, mk); }
    // Montgomery reduction of a 16-limb value held as X + Y*2^32 (< p*2^256): eight SOS steps
    static B200_D void redc16(uint32_t (&X)[16], uint32_t (&Y)[16], uint32_t (&K)[8], uint32_t& c)
    {
      ptx::redc16_step0_fq(X, Y, K, c);
      ptx::redc16_step1_fq(X, Y, K, c);
      ptx::redc16_step2_fq(X, Y, K, c);
      ptx::redc16_step3_fq(X, Y, K, c);
      ptx::redc16_step4_fq(X, Y, K, c);
      ptx::redc16_step5_fq(X, Y, K, c);
      ptx::redc16_step6_fq(X, Y, K, c);
      ptx::redc16_step7_fq(X, Y, K, c);
    }
    static B200_D void addp2(uint32_t (&T)[16]) { ptx::addp2_fq(T); }
#endif
  };

  template <class Cfg>
  struct alignas(16) Fp {
    uint32_t v[8];

    static B200_HD Fp zero()
    {
      Fp r;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = 0;
      return r;
    }
    static B200_HD Fp one()
    {
      Fp r;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = Cfg::ONE(i);
      return r;
    }
    static B200_HD Fp r2()
    {
      Fp r;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = Cfg::R2(i);
      return r;
    }
    static B200_HD Fp raw_one()
    {
      Fp r = zero();
      r.v[0] = 1;
      return r;
    }

    B200_HD bool is_zero() const
    {
      uint32_t o = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o |= v[i];
      return o == 0;
    }
    friend B200_HD bool operator==(const Fp& a, const Fp& b)
    {
      uint32_t o = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o |= a.v[i] ^ b.v[i];
      return o == 0;
    }
    friend B200_HD bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }

    // ------------------------------------------------------------------ add / sub / neg
    friend B200_HD Fp operator+(const Fp& a, const Fp& b)
    {
      Fp r;
#ifdef __CUDA_ARCH__
      uint32_t s[8], t[8];
      ptx::add8(s, a.v, b.v); // a+b < 2p < 2^255: no carry out
      uint32_t bw = Cfg::subp(t, s);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = bw ? s[i] : t[i];
#else
      uint32_t s[8], t[8];
      uint64_t c = 0;
      for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a.v[i] + b.v[i];
        s[i] = (uint32_t)c;
        c >>= 32;
      }
      int64_t bw = 0;
      for (int i = 0; i < 8; ++i) {
        int64_t d = (int64_t)s[i] - Cfg::P(i) + bw;
        t[i] = (uint32_t)d;
        bw = d >> 32;
      }
      for (int i = 0; i < 8; ++i)
        r.v[i] = bw ? s[i] : t[i];
#endif
      return r;
    }

    friend B200_HD Fp operator-(const Fp& a, const Fp& b)
    {
      Fp r;
#ifdef __CUDA_ARCH__
      uint32_t bw = ptx::sub8(r.v, a.v, b.v);
      Cfg::addp_masked(r.v, bw);
#else
      int64_t bw = 0;
      for (int i = 0; i < 8; ++i) {
        int64_t d = (int64_t)a.v[i] - b.v[i] + bw;
        r.v[i] = (uint32_t)d;
        bw = d >> 32;
      }
      if (bw) {
        uint64_t c = 0;
        for (int i = 0; i < 8; ++i) {
          c += (uint64_t)r.v[i] + Cfg::P(i);
          r.v[i] = (uint32_t)c;
          c >>= 32;
        }
      }
#endif
      return r;
    }

    B200_HD Fp neg() const { return zero() - *this; }
    B200_HD Fp dbl() const { return *this + *this; }

    // ------------------------------------------------------------------ Montgomery product
    friend B200_HD Fp operator*(const Fp& a, const Fp& b)
    {
      Fp r;
#ifdef __CUDA_ARCH__
      uint32_t X[8], Y[8];
      ptx::mul_first(X, Y, a.v, b.v[0]);
      Cfg::redc(X, Y);
      ptx::mul_acc(Y, X, a.v, b.v[1]);
      Cfg::redc(Y, X);
      ptx::mul_acc(X, Y, a.v, b.v[2]);
      Cfg::redc(X, Y);
      ptx::mul_acc(Y, X, a.v, b.v[3]);
      Cfg::redc(Y, X);
      ptx::mul_acc(X, Y, a.v, b.v[4]);
      Cfg::redc(X, Y);
      ptx::mul_acc(Y, X, a.v, b.v[5]);
      Cfg::redc(Y, X);
      ptx::mul_acc(X, Y, a.v, b.v[6]);
      Cfg::redc(X, Y);
      ptx::mul_acc(Y, X, a.v, b.v[7]);
      Cfg::redc(Y, X);
      uint32_t s[8], t[8];
      ptx::merge(s, Y, X); // last row's aligned accumulator is Y
      uint32_t bw = Cfg::subp(t, s);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = bw ? s[i] : t[i];
#else
      // host: 4x64-bit CIOS on unsigned __int128 (the prover epilogue's ~2.5k products per proof)
      typedef unsigned __int128 u128;
      uint64_t A[4], B[4], P[4], t[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 4; ++i) {
        A[i] = ((uint64_t)a.v[2 * i + 1] << 32) | a.v[2 * i];
        B[i] = ((uint64_t)b.v[2 * i + 1] << 32) | b.v[2 * i];
        P[i] = ((uint64_t)Cfg::P(2 * i + 1) << 32) | Cfg::P(2 * i);
      }
      // -p^-1 mod 2^64 from the 32-bit constant by one Newton step: x' = x (2 - p x); sign handled below
      uint64_t pinv32 = (uint64_t)(uint32_t)(0u - Cfg::INV); // p^-1 mod 2^32
      uint64_t pinv64 = pinv32 * (2 - P[0] * pinv32);         // p^-1 mod 2^64
      uint64_t ninv = 0 - pinv64;                             // -p^-1 mod 2^64
      for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
          c += (u128)A[j] * B[i] + t[j];
          t[j] = (uint64_t)c;
          c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * ninv;
        c = ((u128)m * P[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) {
          c += (u128)m * P[j] + t[j];
          t[j - 1] = (uint64_t)c;
          c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
      }
      uint64_t u[4];
      unsigned char bw = 0;
      for (int i = 0; i < 4; ++i) {
        u128 d = (u128)t[i] - P[i] - bw;
        u[i] = (uint64_t)d;
        bw = (unsigned char)((d >> 64) & 1);
      }
      bool ge = (t[4] != 0) || (bw == 0);
      for (int i = 0; i < 4; ++i) {
        uint64_t x = ge ? u[i] : t[i];
        r.v[2 * i] = (uint32_t)x;
        r.v[2 * i + 1] = (uint32_t)(x >> 32);
      }
#endif
      return r;
    }

#ifdef __CUDA_ARCH__
    // ---- 512-bit products and their reduction (lazy reduction in the Fq2 tower, dedicated squaring)
    // T = a * b as 16 limbs; operands only need to be < 2^256 (e.g. unreduced sums < 2p)
    static B200_D void mul_wide(uint32_t (&T)[16], const uint32_t (&a)[8], const uint32_t (&b)[8])
    {
      uint32_t X[16], Y[16];
#pragma unroll
      for (int i = 8; i < 16; ++i)
        X[i] = Y[i] = 0;
      ptx::mulwide_row0(X, Y, a, b[0]);
      ptx::mulwide_row1(X, Y, a, b[1]);
      ptx::mulwide_row2(X, Y, a, b[2]);
      ptx::mulwide_row3(X, Y, a, b[3]);
      ptx::mulwide_row4(X, Y, a, b[4]);
      ptx::mulwide_row5(X, Y, a, b[5]);
      ptx::mulwide_row6(X, Y, a, b[6]);
      ptx::mulwide_row7(X, Y, a, b[7]);
      ptx::merge16(T, X, Y);
    }
    // (X + Y*2^32) / 2^256 mod p for a value < p*2^256, fully reduced
    static B200_D Fp redc_xy(uint32_t (&X)[16], uint32_t (&Y)[16])
    {
      uint32_t c = 0, s[8], t[8], K[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        K[i] = 0; // carries leaving a reduction row (the limbs above it hold live data)
      Cfg::redc16(X, Y, K, c);
      ptx::redc16_final(s, X, Y, K, c);
      uint32_t bw = Cfg::subp(t, s);
      Fp r;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r.v[i] = bw ? s[i] : t[i];
      return r;
    }
    static B200_D Fp redc_wide(const uint32_t (&T)[16])
    {
      uint32_t X[16], Y[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        X[i] = T[i];
        Y[i] = 0;
      }
      return redc_xy(X, Y);
    }
    // dedicated squaring: 28 cross products, doubled, + 8 squares (36 wide multiply-adds instead of 64), then reduce.
    // Measured at 60 G/s against 65 G/s for the plain CIOS product (tools/pipes.py): the doubling and the carry
    // words of the SOS reduction cost more ALU/issue slots than the 28 saved multiply-adds return, so sqr() below
    // stays a product; kept for the record and the microbenchmark.
    B200_D Fp sqr_sos() const
    {
      uint32_t X[16], Y[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        X[i] = Y[i] = 0;
      ptx::sqr_cross_row0(X, Y, v);
      ptx::sqr_cross_row1(X, Y, v);
      ptx::sqr_cross_row2(X, Y, v);
      ptx::sqr_cross_row3(X, Y, v);
      ptx::sqr_cross_row4(X, Y, v);
      ptx::sqr_cross_row5(X, Y, v);
      ptx::sqr_cross_row6(X, Y, v);
      ptx::double16(X);
      ptx::double16(Y);
      ptx::sqr_diag(X, v);
      return redc_xy(X, Y);
    }
#endif
    B200_HD Fp sqr() const { return *this * *this; }

    // a*b - c*d (the y-coordinate of every point addition has this shape).  In the base field the single-reduction
    // form (two 512-bit products + one SOS reduction) measured no faster than two CIOS products, so this stays plain;
    // Fq2::mul_sub is where it pays (two reductions instead of four).
    static B200_HD Fp mul_sub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return a * b - c * d; }

    // standard form (as at the reference's API boundary) <-> Montgomery
    static B200_HD Fp to_mont(const Fp& std_form) { return std_form * r2(); }
    static B200_HD Fp from_mont(const Fp& m) { return m * raw_one(); }

    // a^(p-2); inverse(0) == 0 like the reference (modular_arithmetic.h:603)
    B200_HD Fp inverse() const
    {
      // exponent p-2, scanned MSB first
      uint32_t e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        e[i] = Cfg::P(i);
      e[0] -= 2; // P[0] >= 2 for both fields, no borrow
      Fp acc = one();
      for (int i = 7; i >= 0; --i) {
        for (int b = 31; b >= 0; --b) {
          acc = acc.sqr();
          if ((e[i] >> b) & 1) acc = acc * *this;
        }
      }
      return acc;
    }
  };

  typedef Fp<FrCfg> Fr;
  typedef Fp<FqCfg> Fq;

  // ---------------------------------------------------------------------------------------------
  // Fq2 = Fq[u]/(u^2+1)  (nonresidue -1: /root/reference/icicle/include/icicle/fields/snark_fields/bn254_base.h:67-71;
  // reference product: /root/reference/icicle/include/icicle/fields/complex_extension.h:192-219)
  struct alignas(16) Fq2 {
    Fq c0, c1;
    static B200_HD Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static B200_HD Fq2 one() { return {Fq::one(), Fq::zero()}; }
    B200_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    friend B200_HD bool operator==(const Fq2& a, const Fq2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
    friend B200_HD bool operator!=(const Fq2& a, const Fq2& b) { return !(a == b); }
    friend B200_HD Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    friend B200_HD Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    B200_HD Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    B200_HD Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    // Karatsuba: 3 base-field products
    static B200_HD Fq2 mul_inline(const Fq2& a, const Fq2& b)
    {
#ifdef __CUDA_ARCH__
      // lazy reduction: three 512-bit products, two Montgomery reductions (instead of three full products):
      //   c0 = (a0 b0 + p^2 - a1 b1) / R,   c1 = ((a0+a1)(b0+b1) - a0 b0 - a1 b1) / R      (all < p * 2^256)
      uint32_t T0[16], T1[16], T2[16], sa[8], sb[8];
      Fq::mul_wide(T0, a.c0.v, b.c0.v);
      Fq::mul_wide(T1, a.c1.v, b.c1.v);
      ptx::add8(sa, a.c0.v, a.c1.v); // < 2p < 2^255: no reduction needed before the wide product
      ptx::add8(sb, b.c0.v, b.c1.v);
      Fq::mul_wide(T2, sa, sb);
      ptx::sub16(T2, T0);
      ptx::sub16(T2, T1);
      FqCfg::addp2(T0);
      ptx::sub16(T0, T1);
      return {Fq::redc_wide(T0), Fq::redc_wide(T2)};
#else
      Fq t0 = a.c0 * b.c0;
      Fq t1 = a.c1 * b.c1;
      Fq t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
      return {t0 - t1, t2 - t0 - t1};
#endif
    }
    // (c0+c1)(c0-c1), 2 c0 c1 : 2 base-field products
    static B200_HD Fq2 sqr_inline(const Fq2& a)
    {
      Fq s = a.c0 + a.c1, d = a.c0 - a.c1, m = a.c0 * a.c1;
      return {s * d, m.dbl()};
    }
    // Out-of-line on the device: a G2 mixed add is 8 products + 2 squarings in Fq2 = 28 base products; fully inlined
    // that is ~80 KB of SASS, which thrashes the 32 KB instruction cache (ncu before the change:
    // stall_no_instruction 1.3 per issue, fmaheavy 57 % busy). Arguments/result travel in registers.
#if defined(__CUDACC__)
    static __device__ __noinline__ Fq2 mul_call(Fq2 a, Fq2 b) { return mul_inline(a, b); }
    static __device__ __noinline__ Fq2 sqr_call(Fq2 a) { return sqr_inline(a); }
#endif
    friend B200_HD Fq2 operator*(const Fq2& a, const Fq2& b)
    {
#ifdef __CUDA_ARCH__
      return mul_call(a, b);
#else
      return mul_inline(a, b);
#endif
    }
    B200_HD Fq2 sqr() const
    {
#ifdef __CUDA_ARCH__
      return sqr_call(*this);
#else
      return sqr_inline(*this);
#endif
    }
    // a*b - c*d: a fused form (six 512-bit products, two reductions instead of four) was measured SLOWER in the G2
    // accumulate kernel (23.9 vs 22.4 ms for a 3.2 M-point MSM: six live 16-limb arrays spill), so this stays plain
    static B200_HD Fq2 mul_sub(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) { return a * b - c * d; }
    static B200_HD Fq2 to_mont(const Fq2& a) { return {Fq::to_mont(a.c0), Fq::to_mont(a.c1)}; }
    static B200_HD Fq2 from_mont(const Fq2& a) { return {Fq::from_mont(a.c0), Fq::from_mont(a.c1)}; }
    B200_HD Fq2 inverse() const
    {
      Fq n = (c0.sqr() + c1.sqr()).inverse();
      return {c0 * n, (c1 * n).neg()};
    }
  };

} // namespace b200
