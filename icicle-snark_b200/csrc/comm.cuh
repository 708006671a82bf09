// The inter-GPU data plane of the sharded prover and the sharded standalone MSM: NCCL, resolved at
// run time (dlopen "libnccl.so.2": the host process's copy when it already has one - PyTorch's - else the system's),
// so single-GPU users need no NCCL.  One communicator per process (one process per GPU); all collectives are
// enqueued on the library's own streams and ordered with events - the host never waits inside the exchange.
// The reference has no multi-GPU path (device 0 is hard-coded, src/lib.rs:29): this is new surface.
#pragma once
#include <nccl.h>

#include "common.cuh"

struct b200_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

namespace b200 {
  struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
  };
  const NcclApi& nccl(); // resolved once; ok == false when no libnccl.so.2 can be loaded

#define B200_NCCL(call)                                                                                                \
  do {                                                                                                                 \
    ncclResult_t r__ = (call);                                                                                         \
    if (r__ != ncclSuccess) {                                                                                          \
      fprintf(stderr, "[icicle_b200] %s failed: %s (%s:%d)\n", #call, b200::nccl().GetErrorString(r__), __FILE__, __LINE__); \
      return ICICLE_UNKNOWN_FALLBACK_E;                                                                                \
    }                                                                                                                  \
  } while (0)
#define ICICLE_UNKNOWN_FALLBACK_E ((eIcicleError)ICICLE_UNKNOWN_FALLBACK)
} // namespace b200
