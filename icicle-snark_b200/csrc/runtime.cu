// Device runtime behind the icicle_* symbols: a thin, non-throwing CUDA runtime shim.
// Mirrors /root/reference/icicle/src/runtime.cpp:15-277 + backend/cuda/src/cuda_device_api.cu:11-119
// (thread-local active device, allocation tracker, stream-ordered alloc/copy), without the
// multi-backend dispatcher: the only device type is "CUDA".
#include <cstring>
#include <map>
#include <string>

#include <cstdlib>
#include "common.cuh"

namespace b200 {

  static thread_local int tl_device = 0;
  static int g_default_device = 0;
  static thread_local bool tl_device_set = false;

  MemoryTracker& tracker()
  {
    static MemoryTracker t;
    return t;
  }

  int active_device() { return tl_device_set ? tl_device : g_default_device; }

  static std::mutex g_pool_mu;
  static bool g_pool_ready[64] = {false};
  static int g_sm_count[64] = {0};

  eIcicleError ensure_device()
  {
    int dev = active_device();
    B200_CUDA(cudaSetDevice(dev), ICICLE_INVALID_DEVICE);
    if (dev >= 0 && dev < 64 && !g_pool_ready[dev]) {
      std::lock_guard<std::mutex> g(g_pool_mu);
      if (!g_pool_ready[dev]) {
        cudaMemPool_t pool;
        B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev), ICICLE_INVALID_DEVICE);
        uint64_t keep = UINT64_MAX;
        B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep), ICICLE_INVALID_DEVICE);
        // Never let the pool hand stream B a block whose free is still pending on stream A: that inserts a
        // hidden B-after-A dependency and serialises the five concurrent MSM streams of a proof (their scratch
        // blocks are the same size). Reuse only memory whose free has already completed; the pool simply grows
        // to the concurrent peak (a few GB of 180).
        int off = 0;
        B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off), ICICLE_INVALID_DEVICE);
        int n = 0;
        B200_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), ICICLE_INVALID_DEVICE);
        g_sm_count[dev] = n;
        g_pool_ready[dev] = true;
      }
    }
    return ICICLE_SUCCESS;
  }

  int sm_count()
  {
    int dev = active_device();
    if (dev >= 0 && dev < 64 && g_sm_count[dev] > 0) return g_sm_count[dev];
    return 148;
  }

  static bool is_cuda_type(const char* t) { return strncmp(t, "CUDA", 4) == 0; }

} // namespace b200

using namespace b200;

extern "C" {

const char* b200_version(void) { return "icicle-snark-b200 0.1.0 sm_100a"; }

eIcicleError icicle_load_backend(const char*, bool) { return ICICLE_SUCCESS; }
eIcicleError icicle_load_backend_from_env_or_default(void) { return ICICLE_SUCCESS; }

eIcicleError icicle_set_device(const icicleDevice* device)
{
  if (!device || !is_cuda_type(device->type)) return ICICLE_INVALID_DEVICE; // no CPU backend here
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device->id < 0 || device->id >= n) return ICICLE_INVALID_DEVICE;
  tl_device = device->id;
  tl_device_set = true;
  return ensure_device();
}

eIcicleError icicle_set_default_device(const icicleDevice* device)
{
  if (!device || !is_cuda_type(device->type)) return ICICLE_INVALID_DEVICE;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device->id < 0 || device->id >= n) return ICICLE_INVALID_DEVICE;
  g_default_device = device->id;
  return ICICLE_SUCCESS;
}

eIcicleError icicle_get_active_device(icicleDevice* device)
{
  if (!device) return ICICLE_INVALID_POINTER;
  memset(device->type, 0, sizeof(device->type));
  strcpy(device->type, "CUDA");
  device->id = active_device();
  return ICICLE_SUCCESS;
}

eIcicleError icicle_is_host_memory(const void* ptr)
{
  return tracker().identify(ptr) < 0 ? ICICLE_SUCCESS : ICICLE_INVALID_POINTER;
}

eIcicleError icicle_is_active_device_memory(const void* ptr)
{
  int dev = tracker().identify(ptr);
  if (dev < 0) return ICICLE_INVALID_POINTER;
  return dev == active_device() ? ICICLE_SUCCESS : ICICLE_INVALID_POINTER;
}

eIcicleError icicle_get_device_count(int* device_count)
{
  if (!device_count) return ICICLE_INVALID_POINTER;
  B200_CUDA(cudaGetDeviceCount(device_count), ICICLE_INVALID_DEVICE);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_is_device_available(const icicleDevice* device)
{
  if (!device || !is_cuda_type(device->type)) return ICICLE_INVALID_DEVICE;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return ICICLE_INVALID_DEVICE;
  return ICICLE_SUCCESS;
}

eIcicleError icicle_get_registered_devices(char* output, size_t output_size)
{
  const char* s = "CUDA";
  if (!output || output_size < strlen(s) + 1) return ICICLE_INVALID_ARGUMENT;
  strcpy(output, s);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_get_device_properties(icicleDeviceProperties* p)
{
  if (!p) return ICICLE_INVALID_POINTER;
  p->using_host_memory = false;
  p->num_memory_regions = 0;
  p->supports_pinned_memory = true;
  return ICICLE_SUCCESS;
}

eIcicleError icicle_get_available_memory(size_t* total, size_t* free_bytes)
{
  if (!total || !free_bytes) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemGetInfo(free_bytes, total), ICICLE_UNKNOWN_FALLBACK);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_malloc(void** ptr, size_t size)
{
  if (!ptr) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  B200_CUDA(cudaMalloc(ptr, size ? size : 1), ICICLE_ALLOCATION_FAILED);
  tracker().add(*ptr, size ? size : 1, active_device());
  return ICICLE_SUCCESS;
}

eIcicleError icicle_malloc_async(void** ptr, size_t size, icicleStreamHandle stream)
{
  if (!ptr) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  B200_CUDA(cudaMallocAsync(ptr, size ? size : 1, as_stream(stream)), ICICLE_ALLOCATION_FAILED);
  tracker().add(*ptr, size ? size : 1, active_device());
  return ICICLE_SUCCESS;
}

// Like the reference (runtime.cpp:66-93): freeing memory of a non-active device switches, frees, switches back.
eIcicleError icicle_free(void* ptr)
{
  int dev = -1;
  if (!tracker().remove(ptr, &dev)) return ICICLE_INVALID_POINTER;
  int cur = active_device();
  if (dev != cur) cudaSetDevice(dev);
  cudaError_t e = cudaFree(ptr);
  if (dev != cur) cudaSetDevice(cur);
  return translate(e, ICICLE_DEALLOCATION_FAILED);
}

eIcicleError icicle_free_async(void* ptr, icicleStreamHandle stream)
{
  int dev = -1;
  if (!tracker().remove(ptr, &dev)) return ICICLE_INVALID_POINTER;
  if (dev != active_device()) return ICICLE_INVALID_DEVICE;
  return translate(cudaFreeAsync(ptr, as_stream(stream)), ICICLE_DEALLOCATION_FAILED);
}

eIcicleError icicle_memset(void* ptr, int value, size_t size)
{
  if (icicle_is_active_device_memory(ptr) != ICICLE_SUCCESS) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemset(ptr, value, size), ICICLE_UNKNOWN_FALLBACK);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_memset_async(void* ptr, int value, size_t size, icicleStreamHandle stream)
{
  if (icicle_is_active_device_memory(ptr) != ICICLE_SUCCESS) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemsetAsync(ptr, value, size, as_stream(stream)), ICICLE_UNKNOWN_FALLBACK);
  return ICICLE_SUCCESS;
}

// direction from the tracker, as runtime.cpp:152-185 does
static eIcicleError copy_kind(void* dst, const void* src, cudaMemcpyKind* kind)
{
  int d = tracker().identify(dst), s = tracker().identify(src);
  int cur = active_device();
  if (d >= 0 && d != cur) return ICICLE_INVALID_DEVICE;
  if (s >= 0 && s != cur) return ICICLE_INVALID_DEVICE;
  if (d >= 0 && s >= 0)
    *kind = cudaMemcpyDeviceToDevice;
  else if (d >= 0)
    *kind = cudaMemcpyHostToDevice;
  else if (s >= 0)
    *kind = cudaMemcpyDeviceToHost;
  else
    return ICICLE_INVALID_POINTER; // host->host is not a device copy (runtime.cpp:176-180)
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy(void* dst, const void* src, size_t size)
{
  cudaMemcpyKind k;
  B200_TRY(copy_kind(dst, src, &k));
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpy(dst, src, size, k), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy_async(void* dst, const void* src, size_t size, icicleStreamHandle stream)
{
  cudaMemcpyKind k;
  B200_TRY(copy_kind(dst, src, &k));
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpyAsync(dst, src, size, k, as_stream(stream)), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy_to_host(void* dst, const void* src, size_t size)
{
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpy(dst, src, size, cudaMemcpyDeviceToHost), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy_to_host_async(void* dst, const void* src, size_t size, icicleStreamHandle stream)
{
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToHost, as_stream(stream)), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy_to_device(void* dst, const void* src, size_t size)
{
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpy(dst, src, size, cudaMemcpyHostToDevice), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_copy_to_device_async(void* dst, const void* src, size_t size, icicleStreamHandle stream)
{
  B200_TRY(ensure_device());
  B200_CUDA(cudaMemcpyAsync(dst, src, size, cudaMemcpyHostToDevice, as_stream(stream)), ICICLE_COPY_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_create_stream(icicleStreamHandle* stream)
{
  if (!stream) return ICICLE_INVALID_POINTER;
  B200_TRY(ensure_device());
  cudaStream_t s;
  B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), ICICLE_STREAM_CREATION_FAILED);
  *stream = (icicleStreamHandle)s;
  return ICICLE_SUCCESS;
}

eIcicleError icicle_destroy_stream(icicleStreamHandle stream)
{
  B200_CUDA(cudaStreamDestroy(as_stream(stream)), ICICLE_STREAM_DESTRUCTION_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_stream_synchronize(icicleStreamHandle stream)
{
  B200_CUDA(cudaStreamSynchronize(as_stream(stream)), ICICLE_SYNCHRONIZATION_FAILED);
  return ICICLE_SUCCESS;
}

eIcicleError icicle_device_synchronize(void)
{
  B200_TRY(ensure_device());
  B200_CUDA(cudaDeviceSynchronize(), ICICLE_SYNCHRONIZATION_FAILED);
  return ICICLE_SUCCESS;
}

// ------------------------------------------------------------------------------ ConfigExtension
// string-keyed int/bool bag (/root/reference/icicle/include/icicle/config_extension.h); the knobs
// this library reads are listed in DESIGN.md ("c", "precompute", ...). Unknown keys are kept.
struct ConfigExtension {
  std::map<std::string, int> ints;
  std::map<std::string, bool> bools;
};

ConfigExtension* create_config_extension(void) { return new (std::nothrow) ConfigExtension(); }
void destroy_config_extension(ConfigExtension* ext) { delete ext; }
void config_extension_set_int(ConfigExtension* ext, const char* key, int value)
{
  if (ext && key) ext->ints[key] = value;
}
void config_extension_set_bool(ConfigExtension* ext, const char* key, bool value)
{
  if (ext && key) ext->bools[key] = value;
}
int config_extension_get_int(const ConfigExtension* ext, const char* key)
{
  if (!ext || !key) return 0; // the reference throws here; no exception may cross this ABI
  auto it = ext->ints.find(key);
  return it == ext->ints.end() ? 0 : it->second;
}
bool config_extension_get_bool(const ConfigExtension* ext, const char* key)
{
  if (!ext || !key) return false;
  auto it = ext->bools.find(key);
  return it == ext->bools.end() ? false : it->second;
}
ConfigExtension* clone_config_extension(const ConfigExtension* ext)
{
  if (!ext) return nullptr;
  return new (std::nothrow) ConfigExtension(*ext);
}

} // extern "C"
