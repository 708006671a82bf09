/* icicle_b200.h — C ABI of libicicle_b200.so: the B200-native drop-in for the ICICLE symbols that
 * icicle-snark's Groth16 prover binds (BN254 only), plus the fused `b200_groth16_*` fast path.
 *
 * Every declaration cites the reference interface it replaces (paths relative to /root/reference).
 * Struct layouts are byte-identical to the reference's (sizes/offsets in comments were checked with
 * an offsetof probe against the reference headers and are re-checked by tests/test_abi.py against
 * oracle/_ref).  All field elements cross this boundary as 8 x u32 little-endian limbs in
 * STANDARD (non-Montgomery) form unless a config flag says otherwise; G1 affine = {x,y} 64 B,
 * projective = {x,y,z} homogeneous 96 B with identity (0,1,0); G2 doubles every coordinate
 * ([c0][c1]).  No function throws across this boundary; errors come back as eIcicleError codes
 * 0..12 (the Rust mirror, wrappers/rust/icicle-runtime/src/errors.rs:6-20, only knows those).
 *
 * There is no CPU backend behind these symbols: device type "CUDA" (alias "CUDA-B200") only.
 */
#ifndef ICICLE_B200_H
#define ICICLE_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types ------------------------------------------------------------------------------- */
typedef int eIcicleError; /* icicle/include/icicle/errors.h:13-29 */
enum {
  ICICLE_SUCCESS = 0,
  ICICLE_INVALID_DEVICE = 1,
  ICICLE_OUT_OF_MEMORY = 2,
  ICICLE_INVALID_POINTER = 3,
  ICICLE_ALLOCATION_FAILED = 4,
  ICICLE_DEALLOCATION_FAILED = 5,
  ICICLE_COPY_FAILED = 6,
  ICICLE_SYNCHRONIZATION_FAILED = 7,
  ICICLE_STREAM_CREATION_FAILED = 8,
  ICICLE_STREAM_DESTRUCTION_FAILED = 9,
  ICICLE_API_NOT_IMPLEMENTED = 10,
  ICICLE_INVALID_ARGUMENT = 11,
  ICICLE_BACKEND_LOAD_FAILED = 12
};

typedef void* icicleStreamHandle; /* == cudaStream_t; icicle/include/icicle/device_api.h */

typedef struct { uint32_t limbs[8]; } bn254_scalar_t;          /* Fr; math/storage.h:3-47 */
typedef struct { uint32_t limbs[8]; } bn254_fq_t;              /* Fq */
typedef struct { bn254_fq_t x, y; } bn254_affine_t;            /* 64 B;  curves/affine.h */
typedef struct { bn254_fq_t x, y, z; } bn254_projective_t;     /* 96 B;  curves/projective.h:26 */
typedef struct { bn254_fq_t c0, c1; } bn254_fq2_t;             /* fields/complex_extension.h */
typedef struct { bn254_fq2_t x, y; } bn254_g2_affine_t;        /* 128 B */
typedef struct { bn254_fq2_t x, y, z; } bn254_g2_projective_t; /* 192 B */
typedef struct { bn254_fq2_t c[6]; } bn254_fq12_t; /* 384 B: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 of
                                                     fields/snark_fields/bn254_tower.h:22-80 (Fq6 = Fq2[v]/(v^3-9-u),
                                                     Fq12 = Fq6[w]/(w^2-v)), standard form */

typedef struct { char type[64]; int id; } icicleDevice; /* 68 B; icicle/include/icicle/device.h:13-16 */
typedef struct {                                        /* icicle/include/icicle/device_api.h DeviceProperties */
  bool using_host_memory;
  int num_memory_regions;
  bool supports_pinned_memory;
} icicleDeviceProperties;

typedef struct ConfigExtension ConfigExtension; /* opaque; icicle/include/icicle/config_extension.h */

typedef struct {                /* 40 B; icicle/include/icicle/msm.h:21-53 ; rust msm/mod.rs:15-49 */
  icicleStreamHandle stream;    /*  0 */
  int precompute_factor;        /*  8 */
  int c;                        /* 12 */
  int bitsize;                  /* 16 */
  int batch_size;               /* 20 */
  bool are_points_shared_in_batch;  /* 24 */
  bool are_scalars_on_device;       /* 25 */
  bool are_scalars_montgomery_form; /* 26 */
  bool are_points_on_device;        /* 27 */
  bool are_points_montgomery_form;  /* 28 */
  bool are_results_on_device;       /* 29 */
  bool is_async;                    /* 30 */
  ConfigExtension* ext;         /* 32 */
} MSMConfig;

typedef enum { kNN = 0, kNR = 1, kRN = 2, kRR = 3, kNM = 4, kMN = 5 } NTTOrdering; /* ntt.h:37-44 */
typedef enum { kForward = 0, kInverse = 1 } NTTDir;                                /* ntt.h:24-27 */

typedef struct {                /* 64 B; icicle/include/icicle/ntt.h:52-63 ; rust ntt/mod.rs:75-91 */
  icicleStreamHandle stream;    /*  0 */
  bn254_scalar_t coset_gen;     /*  8 */
  int batch_size;               /* 40 */
  bool columns_batch;           /* 44 */
  int ordering;                 /* 48 */
  bool are_inputs_on_device;    /* 52 */
  bool are_outputs_on_device;   /* 53 */
  bool is_async;                /* 54 */
  ConfigExtension* ext;         /* 56 */
} NTTConfig;

typedef struct {                /* 24 B; icicle/include/icicle/ntt.h:91-95 */
  icicleStreamHandle stream;
  bool is_async;
  ConfigExtension* ext;
} NTTInitDomainConfig;

typedef struct {                /* 32 B; icicle/include/icicle/vec_ops.h:17-36 */
  icicleStreamHandle stream;    /*  0 */
  bool is_a_on_device;          /*  8 */
  bool is_b_on_device;          /*  9 */
  bool is_result_on_device;     /* 10 */
  bool is_async;                /* 11 */
  int batch_size;               /* 12 */
  bool columns_batch;           /* 16 */
  ConfigExtension* ext;         /* 24 */
} VecOpsConfig;

/* ---- runtime: icicle/include/icicle/runtime.h:17-281, icicle/src/runtime.cpp:15-386,
 *      rust extern block wrappers/rust/icicle-runtime/src/runtime.rs:10-54.
 *      (C++ references in the reference prototypes are pointers at ABI level.) ------------------ */
eIcicleError icicle_load_backend(const char* path, bool is_recursive);   /* no-op: nothing to dlopen */
eIcicleError icicle_load_backend_from_env_or_default(void);              /* no-op */
eIcicleError icicle_set_device(const icicleDevice* device);              /* thread-local, runtime.cpp:15 */
eIcicleError icicle_set_default_device(const icicleDevice* device);
eIcicleError icicle_get_active_device(icicleDevice* device);
eIcicleError icicle_is_host_memory(const void* ptr);                     /* runtime.cpp:29 */
eIcicleError icicle_is_active_device_memory(const void* ptr);            /* runtime.cpp:35 (range lookup) */
eIcicleError icicle_get_device_count(int* device_count);
eIcicleError icicle_is_device_available(const icicleDevice* device);
eIcicleError icicle_get_registered_devices(char* output, size_t output_size);
eIcicleError icicle_get_device_properties(icicleDeviceProperties* properties);
eIcicleError icicle_get_available_memory(size_t* total, size_t* free_bytes);
eIcicleError icicle_malloc(void** ptr, size_t size);
eIcicleError icicle_malloc_async(void** ptr, size_t size, icicleStreamHandle stream);
eIcicleError icicle_free(void* ptr);
eIcicleError icicle_free_async(void* ptr, icicleStreamHandle stream);
eIcicleError icicle_memset(void* ptr, int value, size_t size);
eIcicleError icicle_memset_async(void* ptr, int value, size_t size, icicleStreamHandle stream);
eIcicleError icicle_copy(void* dst, const void* src, size_t size);
eIcicleError icicle_copy_async(void* dst, const void* src, size_t size, icicleStreamHandle stream);
eIcicleError icicle_copy_to_host(void* dst, const void* src, size_t size);
eIcicleError icicle_copy_to_host_async(void* dst, const void* src, size_t size, icicleStreamHandle stream);
eIcicleError icicle_copy_to_device(void* dst, const void* src, size_t size);
eIcicleError icicle_copy_to_device_async(void* dst, const void* src, size_t size, icicleStreamHandle stream);
eIcicleError icicle_create_stream(icicleStreamHandle* stream);
eIcicleError icicle_destroy_stream(icicleStreamHandle stream);
eIcicleError icicle_stream_synchronize(icicleStreamHandle stream);
eIcicleError icicle_device_synchronize(void);

/* icicle/src/config_extension.cpp:5-38 */
ConfigExtension* create_config_extension(void);
void destroy_config_extension(ConfigExtension* ext);
void config_extension_set_int(ConfigExtension* ext, const char* key, int value);
void config_extension_set_bool(ConfigExtension* ext, const char* key, bool value);
int config_extension_get_int(const ConfigExtension* ext, const char* key);
bool config_extension_get_bool(const ConfigExtension* ext, const char* key);
ConfigExtension* clone_config_extension(const ConfigExtension* ext);

/* ---- MSM: icicle/src/msm.cpp:12-16,28-32,45-49,61-65 -> backend/cuda/src/msm/cuda_msm.cuh:1397-1443 */
eIcicleError bn254_msm(const bn254_scalar_t* scalars, const bn254_affine_t* bases, int msm_size,
                       const MSMConfig* config, bn254_projective_t* results);
eIcicleError bn254_g2_msm(const bn254_scalar_t* scalars, const bn254_g2_affine_t* bases, int msm_size,
                          const MSMConfig* config, bn254_g2_projective_t* results);
eIcicleError bn254_msm_precompute_bases(const bn254_affine_t* input_bases, int bases_size,
                                        const MSMConfig* config, bn254_affine_t* output_bases);
eIcicleError bn254_g2_msm_precompute_bases(const bn254_g2_affine_t* input_bases, int bases_size,
                                           const MSMConfig* config, bn254_g2_affine_t* output_bases);

/* ---- NTT: icicle/src/ntt.cpp:10-14,25-29,40-43,54-63,74-83 -> backend/cuda/include/ntt/ntt.cuh:662-758 */
eIcicleError bn254_ntt(const bn254_scalar_t* input, int size, NTTDir dir, const NTTConfig* config,
                       bn254_scalar_t* output);
eIcicleError bn254_ntt_init_domain(const bn254_scalar_t* primitive_root, const NTTInitDomainConfig* config);
eIcicleError bn254_ntt_release_domain(void);
eIcicleError bn254_get_root_of_unity(uint64_t max_size, bn254_scalar_t* rou);
eIcicleError bn254_get_root_of_unity_from_domain(uint64_t logn, bn254_scalar_t* rou);

/* ---- vec-ops: icicle/src/vec_ops.cpp (add 40, accumulate 54, sub 68-81, mul 84-97, div, sum, product,
 *      scalar_*_vec, convert_montgomery 163-178) -> backend/cuda/src/field/cuda_vec_ops.cu ---------- */
eIcicleError bn254_vector_add(const bn254_scalar_t* a, const bn254_scalar_t* b, uint64_t n,
                              const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_vector_sub(const bn254_scalar_t* a, const bn254_scalar_t* b, uint64_t n,
                              const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_vector_mul(const bn254_scalar_t* a, const bn254_scalar_t* b, uint64_t n,
                              const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_vector_div(const bn254_scalar_t* a, const bn254_scalar_t* b, uint64_t n,
                              const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_vector_accumulate(bn254_scalar_t* a, const bn254_scalar_t* b, uint64_t n,
                                     const VecOpsConfig* config);
eIcicleError bn254_vector_sum(const bn254_scalar_t* a, uint64_t n, const VecOpsConfig* config,
                              bn254_scalar_t* out);
eIcicleError bn254_vector_product(const bn254_scalar_t* a, uint64_t n, const VecOpsConfig* config,
                                  bn254_scalar_t* out);
eIcicleError bn254_scalar_add_vec(const bn254_scalar_t* s, const bn254_scalar_t* v, uint64_t n,
                                  const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_scalar_sub_vec(const bn254_scalar_t* s, const bn254_scalar_t* v, uint64_t n,
                                  const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_scalar_mul_vec(const bn254_scalar_t* s, const bn254_scalar_t* v, uint64_t n,
                                  const VecOpsConfig* config, bn254_scalar_t* out);
eIcicleError bn254_scalar_convert_montgomery(const bn254_scalar_t* in, uint64_t n, bool is_into,
                                             const VecOpsConfig* config, bn254_scalar_t* out);
/* icicle/src/curves/montgomery_conversion.cpp:13-44 */
eIcicleError bn254_affine_convert_montgomery(const bn254_affine_t* in, size_t n, bool is_into,
                                             const VecOpsConfig* config, bn254_affine_t* out);
eIcicleError bn254_projective_convert_montgomery(const bn254_projective_t* in, size_t n, bool is_into,
                                                 const VecOpsConfig* config, bn254_projective_t* out);
eIcicleError bn254_g2_affine_convert_montgomery(const bn254_g2_affine_t* in, size_t n, bool is_into,
                                                const VecOpsConfig* config, bn254_g2_affine_t* out);
eIcicleError bn254_g2_projective_convert_montgomery(const bn254_g2_projective_t* in, size_t n, bool is_into,
                                                    const VecOpsConfig* config, bn254_g2_projective_t* out);

/* ---- host-side scalar helpers (CPU code in the reference's frontend too):
 *      icicle/src/fields/ffi_extern.cpp, icicle/src/curves/ffi_extern.cpp:9-69 (G1), 73-133 (G2) ---- */
void bn254_add(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out);
void bn254_sub(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out);
void bn254_mul(const bn254_scalar_t* a, const bn254_scalar_t* b, bn254_scalar_t* out);
void bn254_inv(const bn254_scalar_t* a, bn254_scalar_t* out);
void bn254_pow(const bn254_scalar_t* base, int exp, bn254_scalar_t* out);
void bn254_from_u32(uint32_t val, bn254_scalar_t* out);
void bn254_generate_scalars(bn254_scalar_t* out, int size);
void bn254_base_field_from_u32(uint32_t val, bn254_fq_t* out);

bool bn254_eq(const bn254_projective_t* a, const bn254_projective_t* b);
bool bn254_is_on_curve(const bn254_projective_t* p);
void bn254_to_affine(const bn254_projective_t* p, bn254_affine_t* out);
void bn254_from_affine(const bn254_affine_t* p, bn254_projective_t* out);
void bn254_generator(bn254_projective_t* out);
void bn254_ecadd(const bn254_projective_t* a, const bn254_projective_t* b, bn254_projective_t* out);
void bn254_ecsub(const bn254_projective_t* a, const bn254_projective_t* b, bn254_projective_t* out);
void bn254_mul_scalar(const bn254_projective_t* p, const bn254_scalar_t* s, bn254_projective_t* out);
void bn254_generate_projective_points(bn254_projective_t* out, int size);
void bn254_generate_affine_points(bn254_affine_t* out, int size);

bool bn254_g2_eq(const bn254_g2_projective_t* a, const bn254_g2_projective_t* b);
bool bn254_g2_is_on_curve(const bn254_g2_projective_t* p);
void bn254_g2_to_affine(const bn254_g2_projective_t* p, bn254_g2_affine_t* out);
void bn254_g2_from_affine(const bn254_g2_affine_t* p, bn254_g2_projective_t* out);
void bn254_g2_generator(bn254_g2_projective_t* out);
void bn254_g2_ecadd(const bn254_g2_projective_t* a, const bn254_g2_projective_t* b, bn254_g2_projective_t* out);
void bn254_g2_ecsub(const bn254_g2_projective_t* a, const bn254_g2_projective_t* b, bn254_g2_projective_t* out);
void bn254_g2_mul_scalar(const bn254_g2_projective_t* p, const bn254_scalar_t* s, bn254_g2_projective_t* out);
void bn254_g2_generate_projective_points(bn254_g2_projective_t* out, int size);
void bn254_g2_generate_affine_points(bn254_g2_affine_t* out, int size);
void bn254_g2_base_field_from_u32(uint32_t val, bn254_fq2_t* out);

/* ---- pairing (CPU code in the reference as well): icicle/src/pairing.cpp:20-24 (rust icicle-core/src/pairing/
 *      mod.rs:37-44); optimal ate, final exponentiation with the hard part of include/icicle/pairing/models/bn.h:66-100;
 *      target-field helpers: icicle/src/fields/ffi_extern_pairing_extension.cpp:6-50 ------------------------------- */
void bn254_pairing(const bn254_affine_t* p, const bn254_g2_affine_t* q, bn254_fq12_t* out);
void bn254_pairing_target_field_generate_scalars(bn254_fq12_t* out, int size);
void bn254_pairing_target_field_sub(bn254_fq12_t* a, bn254_fq12_t* b, bn254_fq12_t* out);
void bn254_pairing_target_field_add(const bn254_fq12_t* a, const bn254_fq12_t* b, bn254_fq12_t* out);
void bn254_pairing_target_field_mul(const bn254_fq12_t* a, const bn254_fq12_t* b, bn254_fq12_t* out);
void bn254_pairing_target_field_inv(const bn254_fq12_t* a, bn254_fq12_t* out);
void bn254_pairing_target_field_pow(const bn254_fq12_t* base, int exp, bn254_fq12_t* out);
void bn254_pairing_target_field_from_u32(uint32_t val, bn254_fq12_t* out);

/* ---- fused Groth16 path (shape B2 of SURVEY 8b): what the Rust groth16_prove body calls instead
 *      of the op-by-op sequence.  Replaces src/cache.rs:117-256 (ZKeyCache), src/proof_helper.rs:31-317
 *      (construct_r1cs + groth16_commitments + epilogue) and src/lib.rs:33-61. -------------------------- */
typedef struct b200_zkey_cache b200_zkey_cache; /* opaque: device-resident, Montgomery-converted zkey */

typedef struct {              /* what Proof{pi_a,pi_b,pi_c} serialises from; proof_helper.rs:22-29 */
  bn254_affine_t pi_a;        /* standard form */
  bn254_g2_affine_t pi_b;
  bn254_affine_t pi_c;
} b200_groth16_proof;

typedef struct {              /* per-phase device times of the last prove, milliseconds (CUDA events) */
  float h2d_ms, r1cs_ms, ntt_ms, msm_g1_ms, msm_g2_ms, total_ms;
} b200_prove_timings;

/* Build the cache from an in-memory .zkey image (the mmap the reference takes in cache.rs:117-181).
 * `precompute` > 1 stores 2^(c*j)-multiples of every base so all MSM windows share one bucket set;
 * 0 = choose automatically (16 when the tables fit in a third of the free device memory, else 1). */
eIcicleError b200_zkey_cache_create(const uint8_t* zkey, size_t zkey_len, int precompute, b200_zkey_cache** out);
eIcicleError b200_zkey_cache_destroy(b200_zkey_cache* cache);
eIcicleError b200_zkey_cache_info(const b200_zkey_cache* cache, uint32_t* n_vars, uint32_t* n_public,
                                  uint32_t* domain_size, uint64_t* n_coef, uint64_t* device_bytes);

/* One proof. `witness` = n_vars x 32 B LE standard form (section 2 of the .wtns, HOST memory);
 * r,s = blinding factors (standard form; NULL => drawn at random like proof_helper.rs:274-285,
 * (1,1) reproduces the reference's `no-randomness` feature, proof_helper.rs:287-295). */
eIcicleError b200_groth16_prove(b200_zkey_cache* cache, const bn254_scalar_t* witness, uint32_t n_witness,
                                const bn254_scalar_t* r, const bn254_scalar_t* s, b200_groth16_proof* proof,
                                b200_prove_timings* timings);

/* Shard hooks for the one-process-per-GPU launch: a rank builds its cache over the
 * contiguous 1/world slice of every base-point section, proves into PARTIAL commitments
 * (A,B1,C,H: projective G1; B2: projective G2 — 4*96+192 = 576 B) that rank 0 folds after one gather. */
typedef struct {
  bn254_projective_t a, b1, c, h;
  bn254_g2_projective_t b2;
} b200_groth16_partials;
eIcicleError b200_zkey_cache_create_sharded(const uint8_t* zkey, size_t zkey_len, int precompute, int rank,
                                            int world, b200_zkey_cache** out);
eIcicleError b200_groth16_commit_partials(b200_zkey_cache* cache, const bn254_scalar_t* witness,
                                          uint32_t n_witness, b200_groth16_partials* out,
                                          b200_prove_timings* timings);
/* Quotient split for N > 1 ranks (optional; commit_partials replicates the chain instead): the three polynomials
 * (0: B.w, 1: A.w, 2: A.w*B.w; src/proof_helper.rs:94-147) are independent, so rank k transforms polynomials
 * [first_poly, first_poly+poly_count) into out_dev (poly_count x domain_size elements, DEVICE memory), the ranks
 * exchange the slices of their H shard (b200_zkey_cache_h_range) with one collective, and commit_end consumes this
 * rank's slice of each (a = A.w', b = B.w', c = product', DEVICE pointers to the shard's first element).
 * commit_begin also uploads the witness and starts the witness-only MSMs, which run during the exchange.
 * The cache is locked from a successful commit_begin until commit_end returns. */
eIcicleError b200_groth16_commit_begin(b200_zkey_cache* cache, const bn254_scalar_t* witness, uint32_t n_witness,
                                       int first_poly, int poly_count, void* out_dev);
eIcicleError b200_groth16_commit_end(b200_zkey_cache* cache, const void* a_dev, const void* b_dev, const void* c_dev,
                                     b200_groth16_partials* out, b200_prove_timings* timings);
eIcicleError b200_zkey_cache_h_range(const b200_zkey_cache* cache, uint32_t* lo, uint32_t* hi);
/* Shard [lo, hi) of an n-element section for `rank` of `world` (host-only). skew = 0: the equal contiguous split used
 * for the H section (and for everything when the quotient chain is replicated). skew > 0: the split used for the
 * signal-indexed sections (A, B1, B2, C) when B200_SHARD_SKEW is set - ranks that transform quotient polynomials get
 * share 1/world + (3/world - polys_owned) * skew, so all ranks finish together. */
eIcicleError b200_shard_range(uint32_t n, int rank, int world, double skew, uint32_t* lo, uint32_t* hi);
/* B1/B2 points kept in this rank's shard after dropping the columns at infinity (signals absent from every B row; dropped
 * when they are >= 1/8 of the shard, B200_SPARSE_B=0/1 overrides) and the shard's signal count. */
eIcicleError b200_zkey_cache_b_points(const b200_zkey_cache* cache, uint32_t* kept, uint32_t* total);
eIcicleError b200_groth16_finish(const b200_zkey_cache* cache, const b200_groth16_partials* parts, int n_parts,
                                 const bn254_scalar_t* r, const bn254_scalar_t* s, b200_groth16_proof* proof);

/* The cut of the five base-point sections (0 = H, 1 = A, 2 = B1, 3 = C, 4 = B2) that rank `rank` of `world` holds
 * (SURVEY 8e).  mode 0 "uniform": every section cut `world` ways (H equally, the signal-indexed ones with `skew`, see
 * b200_shard_range); mode 1 "line": the sections laid end to end, weighted by cost per point, the quotient-polynomial
 * transforms charged to their owners, cut into `world` pieces of equal cost - a rank holds one or two large pieces
 * (whole tables where possible) and keeps the wide Pippenger windows of the single-GPU plan; mode -1: the library
 * default (environment B200_SHARD_PLAN = uniform | line; unset: line for 2 and for >= 6 ranks, uniform otherwise - the
 * measured optimum at 3200k constraints).  b200_zkey_cache_create_sharded cuts
 * with the default mode; b200_zkey_cache_ranges reports what a cache holds. */
eIcicleError b200_shard_plan(uint32_t n_vars, uint32_t domain_size, int rank, int world, int mode, double skew,
                             uint32_t* lo5, uint32_t* hi5);
eIcicleError b200_zkey_cache_ranges(const b200_zkey_cache* cache, uint32_t* lo5, uint32_t* hi5);
int b200_shard_plan_mode(int world); /* the default mode for this world size: 0 uniform, 1 line */

/* ---- multi-GPU data plane inside the library (one process per GPU; NCCL resolved at run time, csrc/comm.cuh) --------
 * The reference has no multi-GPU path (device 0 is hard-coded, src/lib.rs:29); these entry points are what a host in
 * any language binds to shard one proof or one MSM over the GPUs of a box.
 *   b200_comm_unique_id   rank 0 draws the 128-byte rendezvous token; the host hands it to the other ranks
 *   b200_comm_create      collective: joins the communicator on the calling thread's active device
 *   b200_groth16_prove_sharded  one proof over comm->world GPUs: every rank passes the same witness (all ranks host
 *                         memory or all ranks device memory: a host witness is uploaded 1/world per rank and completed by
 *                         one ncclAllGather over NVLink) and a cache built with b200_zkey_cache_create_sharded(rank, world);
 *                         polynomial owners hold their witness-MSM accumulations until their slices are sent; quotient-polynomial
 *                         slices travel by one grouped ncclSend/ncclRecv on the library's streams, the partial sums by
 *                         one 576-byte ncclAllGather; rank 0 folds, blinds and writes `proof` (others may pass NULL)
 *   b200_msm_sharded      every rank holds a contiguous slice of scalars and points (same config semantics as
 *                         bn254_msm / bn254_g2_msm); `result` (HOST memory, projective) holds the total on every rank
 * ICICLE_API_NOT_IMPLEMENTED when no libnccl.so.2 can be loaded. */
typedef struct b200_comm b200_comm;
eIcicleError b200_comm_unique_id(uint8_t* id128);
eIcicleError b200_comm_create(const uint8_t* id128, int rank, int world, b200_comm** out);
eIcicleError b200_comm_destroy(b200_comm* comm);
eIcicleError b200_comm_info(const b200_comm* comm, int* rank, int* world);
eIcicleError b200_groth16_prove_sharded(b200_zkey_cache* cache, b200_comm* comm, const bn254_scalar_t* witness,
                                        uint32_t n_witness, const bn254_scalar_t* r, const bn254_scalar_t* s,
                                        b200_groth16_proof* proof, b200_prove_timings* timings);
eIcicleError b200_msm_sharded(b200_comm* comm, const void* scalars, const void* points, int local_size,
                              const MSMConfig* config, int g2, void* result);

/* File-level mirror of `groth16_prove(witness, zkey, proof, public, device, &mut CacheManager)`
 * (src/lib.rs:33-61): reads .wtns/.zkey, keeps a process-wide cache keyed "{zkey}_{device}",
 * writes proof.json / public.json byte-for-byte as serde_json's pretty printer does. */
eIcicleError b200_groth16_prove_files(const char* witness_path, const char* zkey_path, const char* proof_path,
                                      const char* public_path, const char* device);

/* The proof.json text b200_groth16_prove_files writes for `proof` (serde_json pretty layout, src/proof_helper.rs:308-316);
 * returns its length, 0 if `cap` is too small. Host-only. */
size_t b200_proof_to_json(const b200_groth16_proof* proof, char* out, size_t cap);

/* `groth16_verify_helper` (src/proof_helper.rs:319-372): cpub = IC[0] + sum publics[i] * IC[i+1], then
 * e(-A,B) * e(cpub,gamma_2) * e(C,delta_2) * e(alpha_1,beta_2) == 1 with four pairings on four host threads.
 * Everything in standard form; `ic` holds n_public + 1 points. *valid = 1 / 0. Host-only, like the reference's. */
eIcicleError b200_groth16_verify(const b200_groth16_proof* proof, const bn254_affine_t* vk_alpha_1,
                                 const bn254_g2_affine_t* vk_beta_2, const bn254_g2_affine_t* vk_gamma_2,
                                 const bn254_g2_affine_t* vk_delta_2, const bn254_affine_t* ic,
                                 const bn254_scalar_t* publics, uint64_t n_public, int* valid);
/* File-level mirror of `groth16_verify(proof, public, vk)` (src/lib.rs:63-82): reads snarkjs proof.json, public.json and
 * verification_key.json (src/cache.rs:70-108). INVALID_ARGUMENT for unreadable / malformed files; *valid = 0 where the
 * reference's assert!(pairing_result) would panic. */
eIcicleError b200_groth16_verify_files(const char* proof_path, const char* public_path, const char* vk_path, int* valid);

/* Synthetic-setup / test tool: out[i] = k_i * G1 (64 B affine) or k_i * G2 (128 B affine),
 * k in standard form, host or icicle_malloc'd memory; output Montgomery (as .zkey stores) or standard form.
 * The reference generates its benchmark zkeys with snarkjs (scripts/setup.sh), which is unavailable offline. */
eIcicleError b200_fixed_base_mul(const bn254_scalar_t* k, uint64_t n, int g2, int out_montgomery, void* out);

/* The Pippenger plan derived for an MSM shape (host-only; same code path as bn254_msm and the ZKeyCache):
 * out8 = {c, windows, factor, sets, buckets per set, buckets, work-item cap, n}; hconst9 (optional) = the 288-bit
 * signed-digit recoding constant sum_w 2^(c-1) 2^(c w). c = 0 asks for the heuristic (the reference's is
 * backend/cuda/src/msm/cuda_msm.cuh:45-48). */
eIcicleError b200_msm_plan_info(int n, int c, int bitsize, int precompute_factor, int g2, int32_t* out8, uint32_t* hconst9);
/* kernel launches issued by this library since load (bench.py's gpu_launches) */
unsigned long long b200_launch_count(void);
/* Profiling mode for bench.py's roofline: while enabled, the bucket-accumulation phase (the dominant kernel) of every
 * MSM - standalone or inside a proof - runs isolated (device-wide sync on both sides) and is timed with CUDA events.
 * b200_profile_accumulate(enable) switches it (clearing old records when turning on) and returns the last recorded
 * duration in ms (or -1); b200_profile_records copies the records collected so far, 9 32-bit words each:
 * {g2, tables in the launch, n, windows, c, precompute factor, buckets, batched-affine rounds, float ms}. */
float b200_profile_accumulate(int enable);
int b200_profile_records(int32_t* out9, int cap);

/* library identification: returns a static string "icicle-snark-b200 <version> sm_100a" */
const char* b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ICICLE_B200_H */
