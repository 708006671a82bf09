// Checks inverse_safegcd (csrc/field_inv.cuh) against Fermat's inverse on the host and on the device (research tool,
// lib/libicicle_b200_tools.so; tests/test_host_math.py and tests/test_gpu_ops.py call it).
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "field_inv.cuh"

using namespace b200;

template <class F>
static F sample(uint64_t& s, int k)
{
  F x;
  for (int i = 0; i < 8; ++i) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    x.v[i] = (uint32_t)(s >> 16);
  }
  x.v[7] &= 0x1fffffffu; // < 2^253 < p
  if (k == 0) x = F::zero();
  if (k == 1) x = F::one();
  if (k == 2) x = F::one().neg();
  if (k == 3) { x = F::zero(); x.v[0] = 1; }          // raw 1 = R^-1 in Montgomery terms
  if (k == 4) { x = F::zero(); x.v[7] = 0x10000000u; } // a power of two
  return x;
}

template <class F>
static __global__ void inv_kernel(const F* in, F* out, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = inverse_safegcd(in[i]);
}

template <class F>
static int check(int n, uint64_t seed, int device)
{
  std::vector<F> in(n), got(n);
  uint64_t s = seed | 1;
  for (int i = 0; i < n; ++i)
    in[i] = sample<F>(s, i);
  if (device) {
    F *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&din, n * sizeof(F)) != cudaSuccess || cudaMalloc(&dout, n * sizeof(F)) != cudaSuccess) return -1;
    cudaMemcpy(din, in.data(), n * sizeof(F), cudaMemcpyHostToDevice);
    inv_kernel<F><<<(n + 127) / 128, 128>>>(din, dout, n);
    if (cudaMemcpy(got.data(), dout, n * sizeof(F), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    cudaFree(din);
    cudaFree(dout);
  } else {
    for (int i = 0; i < n; ++i)
      got[i] = inverse_safegcd(in[i]);
  }
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    // x * x^-1 == 1 (or 0 -> 0), and the result is canonical: equal to the Fermat inverse on a few samples
    F prod = in[i] * got[i];
    bool ok = in[i].is_zero() ? got[i].is_zero() : prod == F::one();
    if (ok && i < 16) ok = got[i] == in[i].inverse();
    bad += !ok;
  }
  return bad;
}

extern "C" int b200_inv_check(int n, uint64_t seed, int field_fq, int device)
{
  return field_fq ? check<Fq>(n, seed, device) : check<Fr>(n, seed, device);
}

// host_double_scalar_mul (csrc/host_math.h: the Shamir double-scalar multiplication of the prover's epilogue) against two
// plain double-and-add multiplications, on random points k*G and scalars incl. 0, 1, equal points and opposite points.
// Returns the number of mismatches.
#include "host_math.h"
extern "C" int b200_double_mul_check(int iters, uint64_t seed)
{
  int bad = 0;
  uint64_t s = seed | 1;
  const G1XYZZ g = G1XYZZ::from_affine(g1_generator_mont());
  for (int it = 0; it < iters; ++it) {
    Fr k1 = Fr::from_mont(sample<Fr>(s, -1)), k2 = Fr::from_mont(sample<Fr>(s, -1)), a = Fr::from_mont(sample<Fr>(s, -1)),
       b = Fr::from_mont(sample<Fr>(s, -1));
    if (it == 0) k1 = Fr::zero();
    if (it == 1) { k2 = Fr::zero(); k2.v[0] = 1; }
    G1XYZZ p = host_scalar_mul(g, a), q = host_scalar_mul(g, b);
    if (it == 2) q = p;
    if (it == 3) q = p.neg();
    if (it == 4) q = G1XYZZ::inf();
    G1XYZZ want = host_scalar_mul(p, k1);
    want.add(host_scalar_mul(q, k2));
    G1XYZZ got = host_double_scalar_mul(p, k1, q, k2);
    const bool wi = want.is_inf(), gi = got.is_inf();
    if (wi != gi) {
      ++bad;
    } else if (!wi) {
      Affine<Fq> x = want.to_affine(), y = got.to_affine();
      if (!(x.x == y.x) || !(x.y == y.y)) ++bad;
    }
  }
  return bad;
}
