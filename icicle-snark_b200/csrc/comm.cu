// NCCL resolution + communicator lifetime + the sharded standalone MSM (see comm.cuh).
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "comm.cuh"
#include "curve.cuh"
#include "host_math.h"

namespace b200 {
  const NcclApi& nccl()
  {
    static const NcclApi api = [] {
      NcclApi a;
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return a;
#define B200_SYM(field, name) *(void**)(&a.field) = dlsym(h, name)
      B200_SYM(GetUniqueId, "ncclGetUniqueId");
      B200_SYM(CommInitRank, "ncclCommInitRank");
      B200_SYM(CommDestroy, "ncclCommDestroy");
      B200_SYM(Send, "ncclSend");
      B200_SYM(Recv, "ncclRecv");
      B200_SYM(AllGather, "ncclAllGather");
      B200_SYM(GroupStart, "ncclGroupStart");
      B200_SYM(GroupEnd, "ncclGroupEnd");
      B200_SYM(GetErrorString, "ncclGetErrorString");
#undef B200_SYM
      a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.AllGather && a.GroupStart && a.GroupEnd &&
             a.GetErrorString;
      return a;
    }();
    return api;
  }
} // namespace b200

using namespace b200;

extern "C" {

// rank 0 draws the 128-byte rendezvous token; the host distributes it to the other ranks by whatever channel it has
eIcicleError b200_comm_unique_id(uint8_t* id128)
{
  if (!id128) return ICICLE_INVALID_POINTER;
  if (!nccl().ok) return ICICLE_API_NOT_IMPLEMENTED;
  ncclUniqueId id;
  B200_NCCL(nccl().GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId");
  memcpy(id128, &id, 128);
  return ICICLE_SUCCESS;
}

// collective over all ranks: joins the communicator on the calling thread's active device
eIcicleError b200_comm_create(const uint8_t* id128, int rank, int world, b200_comm** out)
{
  if (!id128 || !out) return ICICLE_INVALID_POINTER;
  if (world < 1 || rank < 0 || rank >= world) return ICICLE_INVALID_ARGUMENT;
  if (!nccl().ok) return ICICLE_API_NOT_IMPLEMENTED;
  B200_TRY(ensure_device());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  b200_comm* c = new b200_comm();
  c->rank = rank;
  c->world = world;
  c->device = active_device();
  ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    fprintf(stderr, "[icicle_b200] ncclCommInitRank failed: %s\n", nccl().GetErrorString(r));
    delete c;
    return (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
  }
  *out = c;
  return ICICLE_SUCCESS;
}

eIcicleError b200_comm_destroy(b200_comm* c)
{
  if (!c) return ICICLE_INVALID_POINTER;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return ICICLE_SUCCESS;
}

eIcicleError b200_comm_info(const b200_comm* c, int* rank, int* world)
{
  if (!c || !rank || !world) return ICICLE_INVALID_POINTER;
  *rank = c->rank;
  *world = c->world;
  return ICICLE_SUCCESS;
}

// Sharded standalone MSM (BASELINE configs[4] at N GPUs): every rank holds a contiguous slice of the scalars and of the
// base points (device or host memory, exactly as for bn254_msm / bn254_g2_msm with the same config), computes its partial
// sum, one ncclAllGather of 96 B (G1) / 192 B (G2) per rank on the config's stream, every rank folds the N partial sums on
// the host and writes the total (projective, standard form) to `result` (HOST memory).  Synchronous.
eIcicleError b200_msm_sharded(b200_comm* c, const void* scalars, const void* points, int local_size, const MSMConfig* cfg_in, int g2, void* result)
{
  if (!c || !cfg_in || !result || !scalars || !points) return ICICLE_INVALID_POINTER;
  if (cfg_in->batch_size > 1) return ICICLE_INVALID_ARGUMENT;
  B200_TRY(ensure_device());
  const size_t psz = g2 ? 192 : 96;
  cudaStream_t st = as_stream(cfg_in->stream);
  uint8_t *d_mine = nullptr, *d_all = nullptr;
  B200_CUDA(cudaMallocAsync((void**)&d_mine, psz, st), ICICLE_ALLOCATION_FAILED);
  B200_CUDA(cudaMallocAsync((void**)&d_all, psz * c->world, st), ICICLE_ALLOCATION_FAILED);
  MSMConfig cfg = *cfg_in;
  cfg.are_results_on_device = true;
  cfg.is_async = true;
  eIcicleError e = g2 ? bn254_g2_msm((const bn254_scalar_t*)scalars, (const bn254_g2_affine_t*)points, local_size, &cfg, (bn254_g2_projective_t*)d_mine)
                      : bn254_msm((const bn254_scalar_t*)scalars, (const bn254_affine_t*)points, local_size, &cfg, (bn254_projective_t*)d_mine);
  std::vector<uint8_t> host(psz * c->world);
  if (e == ICICLE_SUCCESS) {
    ncclResult_t r = nccl().AllGather(d_mine, d_all, psz, ncclUint8, c->comm, st);
    if (r != ncclSuccess) e = (eIcicleError)ICICLE_UNKNOWN_FALLBACK;
  }
  if (e == ICICLE_SUCCESS && cudaMemcpyAsync(host.data(), d_all, host.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) e = ICICLE_COPY_FAILED;
  cudaFreeAsync(d_mine, st);
  cudaFreeAsync(d_all, st);
  if (cudaStreamSynchronize(st) != cudaSuccess && e == ICICLE_SUCCESS) e = ICICLE_SYNCHRONIZATION_FAILED;
  if (e != ICICLE_SUCCESS) return e;
  // fold on the host with the library's own group law (the same helpers as bn254_ecadd)
  if (g2) {
    bn254_g2_projective_t acc;
    memcpy(&acc, host.data(), psz);
    for (int i = 1; i < c->world; ++i) {
      bn254_g2_projective_t p, o;
      memcpy(&p, host.data() + psz * i, psz);
      bn254_g2_ecadd(&acc, &p, &o);
      acc = o;
    }
    memcpy(result, &acc, psz);
  } else {
    bn254_projective_t acc;
    memcpy(&acc, host.data(), psz);
    for (int i = 1; i < c->world; ++i) {
      bn254_projective_t p, o;
      memcpy(&p, host.data() + psz * i, psz);
      bn254_ecadd(&acc, &p, &o);
      acc = o;
    }
    memcpy(result, &acc, psz);
  }
  return ICICLE_SUCCESS;
}

} // extern "C"
