// BN254 G2 MSM instantiation + the bn254_g2_msm / bn254_g2_msm_precompute_bases symbols
// (/root/reference/icicle/src/msm.cpp:28-32,61-65; reference registration cuda_msm_g2.cu:8-10).
#include "msm_impl.cuh"

namespace b200 {
  template eIcicleError msm_enqueue<Fq2>(const MsmPlan&, const Fr*, bool, const Affine<Fq2>*, Projective<Fq2>*, cudaStream_t);
  template eIcicleError msm_reduce_enqueue<Fq2>(const MsmPlan&, const MsmSorted&, const Affine<Fq2>* const*, int, Projective<Fq2>*, cudaStream_t, cudaEvent_t);
  template eIcicleError precompute_enqueue<Fq2>(const Affine<Fq2>*, bool, int, int, int, Affine<Fq2>*, bool, cudaStream_t);
} // namespace b200

using namespace b200;

extern "C" {

eIcicleError bn254_g2_msm(
  const bn254_scalar_t* scalars, const bn254_g2_affine_t* bases, int msm_size, const MSMConfig* config,
  bn254_g2_projective_t* results)
{
  return msm_api<Fq2>(scalars, bases, msm_size, config, results, true);
}

eIcicleError bn254_g2_msm_precompute_bases(
  const bn254_g2_affine_t* input_bases, int bases_size, const MSMConfig* config, bn254_g2_affine_t* output_bases)
{
  return precompute_api<Fq2>(input_bases, bases_size, config, output_bases, true);
}

} // extern "C"
