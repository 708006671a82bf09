// Host-side scalar/point helpers of the C ABI: bn254_{add,sub,mul,inv,pow,...}, bn254_{ecadd,ecsub,
// mul_scalar,to_affine,eq,...} and their G2 twins.  In the reference these are CPU code compiled
// into the frontend as well (/root/reference/icicle/src/fields/ffi_extern.cpp,
// /root/reference/icicle/src/curves/ffi_extern.cpp:9-69,73-133); the prover's epilogue
// (src/proof_helper.rs:280-283) is their only hot-path user (~20 group ops per proof).
// Same field/curve templates as the device code, host instantiation (portable 64-bit CIOS).
#include <random>
#include <vector>

#include "common.cuh"
#include "curve.cuh"
#include "host_math.h"
#include <cerrno>
#include <sys/random.h>

namespace b200 {

  template <class F>
  static bool proj_eq(const Projective<F>& a, const Projective<F>& b)
  {
    // cross-multiplied equality; (0,0,0) equals nothing (ffi_extern.cpp:9-16, projective.h:212-215)
    bool a0 = a.x.is_zero() && a.y.is_zero() && a.z.is_zero();
    bool b0 = b.x.is_zero() && b.y.is_zero() && b.z.is_zero();
    if (a0 || b0) return false;
    return (a.x * b.z == b.x * a.z) && (a.y * b.z == b.y * a.z);
  }

  template <>
  Fq curve_b<Fq>()
  {
    Fq b = Fq::zero();
    b.v[0] = 3;
    return Fq::to_mont(b);
  }
  template <>
  Fq2 curve_b<Fq2>()
  {
    // b' = 3/(9+u)  (/root/reference/icicle/include/icicle/curves/params/bn254.h:41-44)
    static const uint32_t re[8] = {0x24a138e5, 0x3267e6dc, 0x59dbefa3, 0xb5b4c5e5, 0x1be06ac3, 0x81be1899, 0xceb8aaae, 0x2b149d40};
    static const uint32_t im[8] = {0x85c315d2, 0xe4a2bd06, 0xe52d1852, 0xa74fa084, 0xeed8fdf4, 0xcd2cafad, 0x3af0fed4, 0x009713b0};
    Fq2 b;
    memcpy(b.c0.v, re, 32);
    memcpy(b.c1.v, im, 32);
    return Fq2::to_mont(b);
  }

  template <class F>
  static bool proj_on_curve(const Projective<F>& p)
  {
    // y^2 z = x^3 + b z^3 ; identity (0,y,0) is on the curve
    if (p.z.is_zero()) return p.x.is_zero() && !p.y.is_zero();
    F z2 = p.z.sqr();
    return p.y.sqr() * p.z == p.x.sqr() * p.x + curve_b<F>() * z2 * p.z;
  }

  G1Affine g1_generator_mont()
  {
    Fq x = Fq::zero(), y = Fq::zero();
    x.v[0] = 1;
    y.v[0] = 2;
    return {Fq::to_mont(x), Fq::to_mont(y)};
  }

  G2Affine g2_generator_mont()
  {
    // /root/reference/icicle/include/icicle/curves/params/bn254.h:32-39
    static const uint32_t xr[8] = {0xd992f6ed, 0x46debd5c, 0xf75edadd, 0x674322d4, 0x5e5c4479, 0x426a0066, 0x121f1e76, 0x1800deef};
    static const uint32_t xi[8] = {0xaef312c2, 0x97e485b7, 0x35a9e712, 0xf1aa4933, 0x31fb5d25, 0x7260bfb7, 0x920d483a, 0x198e9393};
    static const uint32_t yr[8] = {0x66fa7daa, 0x4ce6cc01, 0x0c43d37b, 0xe3d1e769, 0x8dcb408f, 0x4aab7180, 0xdb8c6deb, 0x12c85ea5};
    static const uint32_t yi[8] = {0xd122975b, 0x55acdadc, 0x70b38ef3, 0xbc4b3133, 0x690c3395, 0xec9e99ad, 0x585ff075, 0x090689d0};
    G2Affine g;
    memcpy(g.x.c0.v, xr, 32);
    memcpy(g.x.c1.v, xi, 32);
    memcpy(g.y.c0.v, yr, 32);
    memcpy(g.y.c1.v, yi, 32);
    return {Fq2::to_mont(g.x), Fq2::to_mont(g.y)};
  }

  // Blinding factors must come from the kernel's CSPRNG: a 32-bit-seeded mt19937 (what the reference's
  // ScalarCfg::generate_random amounts to) lets an attacker enumerate seeds and confirm a candidate witness from pi_a.
  bool host_secure_random_fr(Fr& out)
  {
    for (int tries = 0; tries < 64; ++tries) {
      Fr x;
      size_t got = 0;
      while (got < sizeof(x.v)) {
        ssize_t r = getrandom(reinterpret_cast<uint8_t*>(x.v) + got, sizeof(x.v) - got, 0);
        if (r < 0) {
          if (errno == EINTR) continue;
          return false;
        }
        got += (size_t)r;
      }
      x.v[7] &= 0x3fffffff; // rejection sampling on 254 bits: uniform in [0, r)
      for (int i = 7; i >= 0; --i) {
        if (x.v[i] != FrCfg::P(i)) {
          if (x.v[i] < FrCfg::P(i)) {
            out = x;
            return true;
          }
          break;
        }
      }
    }
    return false;
  }

  Fr host_random_fr(std::mt19937_64& rng)
  {
    // uniform in [0, r) by rejection on 254 bits; returned in STANDARD form
    for (;;) {
      Fr x;
      for (int i = 0; i < 8; i += 2) {
        uint64_t w = rng();
        x.v[i] = (uint32_t)w;
        x.v[i + 1] = (uint32_t)(w >> 32);
      }
      x.v[7] &= 0x3fffffff;
      bool lt = false;
      for (int i = 7; i >= 0; --i) {
        if (x.v[i] != FrCfg::P(i)) {
          lt = x.v[i] < FrCfg::P(i);
          break;
        }
      }
      if (lt) return x;
    }
  }

  template <class F>
  static void batch_to_affine(const std::vector<XYZZ<F>>& in, Affine<F>* out_mont)
  {
    // Montgomery's trick over d_i = ZZ_i*ZZZ_i
    size_t n = in.size();
    std::vector<F> d(n), pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) {
      d[i] = in[i].is_inf() ? F::one() : in[i].zz * in[i].zzz;
      pre[i] = acc;
      acc = acc * d[i];
    }
    F inv = acc.inverse();
    for (size_t i = n; i-- > 0;) {
      F di = inv * pre[i];
      inv = inv * d[i];
      if (in[i].is_inf())
        out_mont[i] = Affine<F>::inf();
      else
        out_mont[i] = {in[i].x * in[i].zzz * di, in[i].y * in[i].zz * di};
    }
  }

  template <class F>
  static void gen_affine_points(Affine<F>* out_std, int size, const Affine<F>& gen)
  {
    if (size <= 0) return;
    std::mt19937_64 rng(std::random_device{}());
    std::vector<XYZZ<F>> pts(size);
    XYZZ<F> cur = host_scalar_mul(XYZZ<F>::from_affine(gen), host_random_fr(rng));
    XYZZ<F> step = host_scalar_mul(XYZZ<F>::from_affine(gen), host_random_fr(rng));
    for (int i = 0; i < size; ++i) {
      pts[i] = cur;
      cur.add(step);
    }
    batch_to_affine(pts, out_std);
    for (int i = 0; i < size; ++i)
      out_std[i] = affine_from_mont(out_std[i]);
  }

} // namespace b200

using namespace b200;

template <class T, class U>
static inline T& as(U* p)
{
  return *reinterpret_cast<T*>(p);
}
template <class T, class U>
static inline const T& as(const U* p)
{
  return *reinterpret_cast<const T*>(p);
}

#define EXPORT __attribute__((visibility("default")))

extern "C" {

// ---------------------------------------------------------------------------------- Fr
EXPORT void bn254_add(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out)
{
  as<Fr>(out) = as<Fr>(a) + as<Fr>(b);
}
EXPORT void bn254_sub(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out)
{
  as<Fr>(out) = as<Fr>(a) - as<Fr>(b);
}
EXPORT void bn254_mul(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out)
{
  as<Fr>(out) = (as<Fr>(a) * as<Fr>(b)) * Fr::r2();
}
EXPORT void bn254_inv(const bn254_scalar_t* a, bn254_scalar_t* out)
{
  as<Fr>(out) = Fr::from_mont(Fr::to_mont(as<Fr>(a)).inverse());
}
EXPORT void bn254_pow(const bn254_scalar_t* base, int exp, bn254_scalar_t* out)
{
  Fr b = Fr::to_mont(as<Fr>(base)), acc = Fr::one();
  unsigned e = (unsigned)exp;
  while (e) {
    if (e & 1) acc = acc * b;
    b = b.sqr();
    e >>= 1;
  }
  as<Fr>(out) = Fr::from_mont(acc);
}
EXPORT void bn254_from_u32(uint32_t val, bn254_scalar_t* out)
{
  Fr r = Fr::zero();
  r.v[0] = val;
  as<Fr>(out) = r;
}
EXPORT void bn254_generate_scalars(bn254_scalar_t* out, int size)
{
  std::mt19937_64 rng(std::random_device{}());
  for (int i = 0; i < size; ++i)
    as<Fr>(out + i) = host_random_fr(rng);
}
EXPORT void bn254_base_field_from_u32(uint32_t val, bn254_fq_t* out)
{
  Fq r = Fq::zero();
  r.v[0] = val;
  as<Fq>(out) = r;
}
EXPORT void bn254_g2_base_field_from_u32(uint32_t val, bn254_fq2_t* out)
{
  Fq2 r = Fq2::zero();
  r.c0.v[0] = val;
  as<Fq2>(out) = r;
}

// ---------------------------------------------------------------------------------- G1 / G2
#define CURVE_FFI(PFX, F, AFF_T, PROJ_T, GEN)                                                                          \
  EXPORT bool PFX##eq(const PROJ_T* a, const PROJ_T* b)                                                                \
  {                                                                                                                    \
    return proj_eq(proj_to_mont(as<Projective<F>>(a)), proj_to_mont(as<Projective<F>>(b)));                            \
  }                                                                                                                    \
  EXPORT bool PFX##is_on_curve(const PROJ_T* p) { return proj_on_curve(proj_to_mont(as<Projective<F>>(p))); }          \
  EXPORT void PFX##to_affine(const PROJ_T* p, AFF_T* out)                                                              \
  {                                                                                                                    \
    as<Affine<F>>(out) = affine_from_mont(xyzz_from_projective(proj_to_mont(as<Projective<F>>(p))).to_affine());      \
  }                                                                                                                    \
  EXPORT void PFX##from_affine(const AFF_T* p, PROJ_T* out)                                                            \
  {                                                                                                                    \
    const Affine<F>& a = as<Affine<F>>(p);                                                                             \
    Projective<F> r;                                                                                                   \
    if (a.is_inf()) {                                                                                                  \
      r = {F::zero(), F::from_mont(F::one()), F::zero()};                                                              \
    } else {                                                                                                           \
      r = {a.x, a.y, F::from_mont(F::one())};                                                                          \
    }                                                                                                                  \
    as<Projective<F>>(out) = r;                                                                                        \
  }                                                                                                                    \
  EXPORT void PFX##generator(PROJ_T* out)                                                                              \
  {                                                                                                                    \
    as<Projective<F>>(out) = proj_from_mont(XYZZ<F>::from_affine(GEN()).to_projective());                              \
  }                                                                                                                    \
  EXPORT void PFX##ecadd(const PROJ_T* a, const PROJ_T* b, PROJ_T* out)                                                \
  {                                                                                                                    \
    XYZZ<F> x = xyzz_from_projective(proj_to_mont(as<Projective<F>>(a)));                                              \
    x.add(xyzz_from_projective(proj_to_mont(as<Projective<F>>(b))));                                                   \
    as<Projective<F>>(out) = proj_from_mont(x.to_projective());                                                        \
  }                                                                                                                    \
  EXPORT void PFX##ecsub(const PROJ_T* a, const PROJ_T* b, PROJ_T* out)                                                \
  {                                                                                                                    \
    XYZZ<F> x = xyzz_from_projective(proj_to_mont(as<Projective<F>>(a)));                                              \
    x.add(xyzz_from_projective(proj_to_mont(as<Projective<F>>(b))).neg());                                             \
    as<Projective<F>>(out) = proj_from_mont(x.to_projective());                                                        \
  }                                                                                                                    \
  EXPORT void PFX##mul_scalar(const PROJ_T* p, const bn254_scalar_t* s, PROJ_T* out)                                   \
  {                                                                                                                    \
    XYZZ<F> x = xyzz_from_projective(proj_to_mont(as<Projective<F>>(p)));                                              \
    as<Projective<F>>(out) = proj_from_mont(host_scalar_mul(x, as<Fr>(s)).to_projective());                            \
  }                                                                                                                    \
  EXPORT void PFX##generate_affine_points(AFF_T* out, int size)                                                        \
  {                                                                                                                    \
    gen_affine_points<F>(reinterpret_cast<Affine<F>*>(out), size, GEN());                                              \
  }                                                                                                                    \
  EXPORT void PFX##generate_projective_points(PROJ_T* out, int size)                                                   \
  {                                                                                                                    \
    std::vector<Affine<F>> tmp(size > 0 ? size : 0);                                                                   \
    gen_affine_points<F>(tmp.data(), size, GEN());                                                                     \
    for (int i = 0; i < size; ++i)                                                                                     \
      as<Projective<F>>(out + i) = {tmp[i].x, tmp[i].y, F::from_mont(F::one())};                                       \
  }

CURVE_FFI(bn254_, Fq, bn254_affine_t, bn254_projective_t, g1_generator_mont)
CURVE_FFI(bn254_g2_, Fq2, bn254_g2_affine_t, bn254_g2_projective_t, g2_generator_mont)

} // extern "C"
