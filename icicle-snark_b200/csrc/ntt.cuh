// Fr number-theoretic transform for BN254, sm_100a.
//
// Replaces /root/reference/icicle/backend/cuda/include/ntt/ntt.cuh:394-579 (Domain), :662-758 (ntt_cuda)
// and /root/reference/icicle/backend/cuda/src/ntt/mixed_radix_ntt.cu:755-1017 (large_ntt, mixed_radix_ntt).
//
//   reference                                          here
//   ---------                                          ----
//   in-place DIT/DIF stages + a separate digit-        Stockham autosort passes (out-of-place ping-pong): every
//   reversal pass (cycle-following when in place,      pass reads natural order and writes natural order, so
//   mixed_radix_ntt.cu:60-126)                         kNN needs no permutation pass at all
//   64-thread CTAs, radix 16/32/64 per launch,         one CTA = 2^k points x 2^m adjacent columns staged in
//   4 launches at 2^22                                 shared memory, k <= 8: 3 launches at 2^22
//   5 twiddle tables (basic/internal/external)         one table w^i, i in [0, N_max], Montgomery form, resident
//                                                      in HBM; the CTA's 2^(k-1) internal twiddles are staged
//                                                      in shared memory
//   separate normalise / coset kernels                 1/N and an optional per-index table (the prover's coset
//                                                      shift) are fused into the last pass's store
//
// The kernels are agnostic to the data's form: a Montgomery product of a standard-form value with a
// Montgomery-form twiddle is the standard-form product, so API callers (standard form) and the fused
// prover (Montgomery form) share the same code with no conversion.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace b200 {

  struct NttDomain {
    int max_log = -1;      // table holds w^i for i in [0, 2^max_log]
    Fr* tw = nullptr;      // device, Montgomery form
    Fr root_std;           // the primitive root passed to init (standard form)
  };

  // device-current domain (nullptr if not initialised)
  const NttDomain* ntt_domain();

  // Enqueue `batch` transforms of size 2^logn. in/out are device pointers (may alias).
  // post_table (nullable): out[i] *= post_table[i] (natural output index), fused into the last pass.
  // The inverse transform includes the 1/N scaling.
  eIcicleError ntt_enqueue(
    const Fr* in, Fr* out, int logn, bool inverse, int batch, bool columns_batch, const Fr* post_table, cudaStream_t st);

  eIcicleError ntt_init_domain_host(const Fr& root_std, cudaStream_t st);

  Fr host_omega(int logn); // w of order 2^logn, standard form (bn254_scalar.h:68-69, modular_arithmetic.h:61-73)

} // namespace b200
