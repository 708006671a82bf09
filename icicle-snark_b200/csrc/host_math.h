// Host-side group helpers shared by host_math.cu (FFI scalar helpers) and the prover epilogue.
#pragma once
#include <random>

#include "curve.cuh"

namespace b200 {

  template <class F>
  inline Projective<F> proj_to_mont(const Projective<F>& p)
  {
    return {F::to_mont(p.x), F::to_mont(p.y), F::to_mont(p.z)};
  }
  template <class F>
  inline Projective<F> proj_from_mont(const Projective<F>& p)
  {
    return {F::from_mont(p.x), F::from_mont(p.y), F::from_mont(p.z)};
  }
  template <class F>
  inline Affine<F> affine_to_mont(const Affine<F>& p)
  {
    return {F::to_mont(p.x), F::to_mont(p.y)};
  }
  template <class F>
  inline Affine<F> affine_from_mont(const Affine<F>& p)
  {
    return {F::from_mont(p.x), F::from_mont(p.y)};
  }

  // k*P, k in STANDARD form; plain MSB-first double-and-add (host, a handful of calls per proof)
  template <class F>
  inline XYZZ<F> host_scalar_mul(const XYZZ<F>& p, const Fr& k_std)
  {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = 7; i >= 0; --i) {
      for (int b = 31; b >= 0; --b) {
        acc = acc.dbl();
        if ((k_std.v[i] >> b) & 1) acc.add(p);
      }
    }
    return acc;
  }

  // k1*P + k2*Q with one shared doubling chain (Shamir's trick): 256 doublings + ~192 additions instead of 512 + 256.
  // The prover's epilogue needs s*pi_a + r*pi_b1 after the GPU is done - the only host arithmetic on the critical path.
  template <class F>
  inline XYZZ<F> host_double_scalar_mul(const XYZZ<F>& p, const Fr& k1_std, const XYZZ<F>& q, const Fr& k2_std)
  {
    XYZZ<F> pq = p;
    pq.add(q);
    const XYZZ<F>* tab[4] = {nullptr, &p, &q, &pq};
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = 7; i >= 0; --i) {
      for (int b = 31; b >= 0; --b) {
        acc = acc.dbl();
        const int sel = (int)((k1_std.v[i] >> b) & 1) | ((int)((k2_std.v[i] >> b) & 1) << 1);
        if (sel) acc.add(*tab[sel]);
      }
    }
    return acc;
  }

  // curve coefficient b (G1: 3, G2: 3/(9+u)) in Montgomery form; specialised in host_math.cu
  template <class F>
  F curve_b();
  template <>
  Fq curve_b<Fq>();
  template <>
  Fq2 curve_b<Fq2>();

  G1Affine g1_generator_mont();
  G2Affine g2_generator_mont();
  Fr host_random_fr(std::mt19937_64& rng); // standard form, uniform in [0, r): test-data generators only
  bool host_secure_random_fr(Fr& out);     // getrandom(2) + rejection sampling; false if the kernel CSPRNG fails

} // namespace b200
