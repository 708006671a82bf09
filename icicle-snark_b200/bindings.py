"""ctypes view of the C ABI declared in include/icicle_b200.h.

Host-side mirror of the reference's Rust FFI layer for the Groth16 path
(/root/reference/wrappers/rust/icicle-runtime/src/runtime.rs:10-54,
 /root/reference/wrappers/rust/icicle-core/src/msm/mod.rs:13-49,106-154,193-259,
 /root/reference/wrappers/rust/icicle-core/src/ntt/mod.rs:73-107,202-216,311-355,
 /root/reference/wrappers/rust/icicle-core/src/vec_ops/mod.rs:6-32,219-245):
same config structs (field order and defaults), same entry-point names, same
error behaviour (every call returns an eIcicleError; `check` raises like the
Rust `.unwrap()` panics).

The struct layouts are the ABI, so the very same classes also drive the
reference's own CPU library in the tests (oracle/ref_cpu.py) - that is the
drop-in claim, exercised.  No arithmetic happens in this file.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SCALAR_WORDS = 8
G1_AFFINE_WORDS, G1_PROJ_WORDS = 16, 24
G2_AFFINE_WORDS, G2_PROJ_WORDS = 32, 48
FQ12_WORDS = 96

ERRORS = [
    "SUCCESS", "INVALID_DEVICE", "OUT_OF_MEMORY", "INVALID_POINTER", "ALLOCATION_FAILED",
    "DEALLOCATION_FAILED", "COPY_FAILED", "SYNCHRONIZATION_FAILED", "STREAM_CREATION_FAILED",
    "STREAM_DESTRUCTION_FAILED", "API_NOT_IMPLEMENTED", "INVALID_ARGUMENT", "BACKEND_LOAD_FAILED",
]

kNN, kNR, kRN, kRR, kNM, kMN = range(6)
kForward, kInverse = 0, 1


class IcicleError(RuntimeError):
    def __init__(self, code, what):
        self.code = code
        name = ERRORS[code] if 0 <= code < len(ERRORS) else "UNKNOWN_ERROR"
        super().__init__(f"{what}: eIcicleError {code} ({name})")


def check(code, what="icicle call"):
    if code != 0:
        raise IcicleError(code, what)


class Device(C.Structure):  # icicle/include/icicle/device.h:13-16 (68 B)
    _fields_ = [("type", C.c_char * 64), ("id", C.c_int)]

    @classmethod
    def new(cls, type_="CUDA", id_=0):
        d = cls()
        d.type = type_.encode()
        d.id = id_
        return d


class MSMConfig(C.Structure):  # icicle/include/icicle/msm.h:21-53 (40 B)
    _fields_ = [
        ("stream", C.c_void_p), ("precompute_factor", C.c_int), ("c", C.c_int), ("bitsize", C.c_int),
        ("batch_size", C.c_int), ("are_points_shared_in_batch", C.c_bool), ("are_scalars_on_device", C.c_bool),
        ("are_scalars_montgomery_form", C.c_bool), ("are_points_on_device", C.c_bool),
        ("are_points_montgomery_form", C.c_bool), ("are_results_on_device", C.c_bool), ("is_async", C.c_bool),
        ("ext", C.c_void_p),
    ]

    @classmethod
    def default(cls):  # rust msm/mod.rs:33-49
        return cls(None, 1, 0, 0, 1, True, False, False, False, False, False, False, None)


class NTTConfig(C.Structure):  # icicle/include/icicle/ntt.h:52-63 (64 B)
    _fields_ = [
        ("stream", C.c_void_p), ("coset_gen", C.c_uint32 * 8), ("batch_size", C.c_int), ("columns_batch", C.c_bool),
        ("ordering", C.c_int), ("are_inputs_on_device", C.c_bool), ("are_outputs_on_device", C.c_bool),
        ("is_async", C.c_bool), ("ext", C.c_void_p),
    ]

    @classmethod
    def default(cls):  # rust ntt/mod.rs:93-107: coset_gen = one
        one = (C.c_uint32 * 8)(1, 0, 0, 0, 0, 0, 0, 0)
        return cls(None, one, 1, False, kNN, False, False, False, None)


class NTTInitDomainConfig(C.Structure):  # ntt.h:91-95 (24 B)
    _fields_ = [("stream", C.c_void_p), ("is_async", C.c_bool), ("ext", C.c_void_p)]


class VecOpsConfig(C.Structure):  # icicle/include/icicle/vec_ops.h:17-36 (32 B)
    _fields_ = [
        ("stream", C.c_void_p), ("is_a_on_device", C.c_bool), ("is_b_on_device", C.c_bool),
        ("is_result_on_device", C.c_bool), ("is_async", C.c_bool), ("batch_size", C.c_int),
        ("columns_batch", C.c_bool), ("ext", C.c_void_p),
    ]

    @classmethod
    def default(cls):
        return cls(None, False, False, False, False, 1, False, None)


class Groth16Proof(C.Structure):  # include/icicle_b200.h b200_groth16_proof (standard form)
    _fields_ = [("pi_a", C.c_uint32 * 16), ("pi_b", C.c_uint32 * 32), ("pi_c", C.c_uint32 * 16)]


class ProveTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "r1cs_ms", "ntt_ms", "msm_g1_ms", "msm_g2_ms", "total_ms")]


class Groth16Partials(C.Structure):
    _fields_ = [("a", C.c_uint32 * 24), ("b1", C.c_uint32 * 24), ("c", C.c_uint32 * 24), ("h", C.c_uint32 * 24),
                ("b2", C.c_uint32 * 48)]


assert C.sizeof(Device) == 68 and C.sizeof(MSMConfig) == 40 and C.sizeof(NTTConfig) == 64
assert C.sizeof(NTTInitDomainConfig) == 24 and C.sizeof(VecOpsConfig) == 32
assert MSMConfig.ext.offset == 32 and NTTConfig.batch_size.offset == 40 and NTTConfig.ordering.offset == 48
assert NTTConfig.ext.offset == 56 and VecOpsConfig.batch_size.offset == 12 and VecOpsConfig.ext.offset == 24

# every symbol include/icicle_b200.h declares that is common to the reference's libraries
ABI_SYMBOLS = """
icicle_load_backend icicle_load_backend_from_env_or_default icicle_set_device icicle_set_default_device
icicle_get_active_device icicle_is_host_memory icicle_is_active_device_memory icicle_get_device_count
icicle_is_device_available icicle_get_registered_devices icicle_get_device_properties icicle_get_available_memory
icicle_malloc icicle_malloc_async icicle_free icicle_free_async icicle_memset icicle_memset_async icicle_copy
icicle_copy_async icicle_copy_to_host icicle_copy_to_host_async icicle_copy_to_device icicle_copy_to_device_async
icicle_create_stream icicle_destroy_stream icicle_stream_synchronize icicle_device_synchronize
create_config_extension destroy_config_extension config_extension_set_int config_extension_set_bool
config_extension_get_int config_extension_get_bool clone_config_extension
bn254_msm bn254_g2_msm bn254_msm_precompute_bases bn254_g2_msm_precompute_bases
bn254_ntt bn254_ntt_init_domain bn254_ntt_release_domain bn254_get_root_of_unity bn254_get_root_of_unity_from_domain
bn254_vector_add bn254_vector_sub bn254_vector_mul bn254_vector_div bn254_vector_accumulate bn254_vector_sum
bn254_vector_product bn254_scalar_add_vec bn254_scalar_sub_vec bn254_scalar_mul_vec bn254_scalar_convert_montgomery
bn254_affine_convert_montgomery bn254_projective_convert_montgomery bn254_g2_affine_convert_montgomery
bn254_g2_projective_convert_montgomery
bn254_add bn254_sub bn254_mul bn254_inv bn254_pow bn254_from_u32 bn254_generate_scalars bn254_base_field_from_u32
bn254_eq bn254_is_on_curve bn254_to_affine bn254_from_affine bn254_generator bn254_ecadd bn254_ecsub bn254_mul_scalar
bn254_generate_projective_points bn254_generate_affine_points
bn254_g2_eq bn254_g2_is_on_curve bn254_g2_to_affine bn254_g2_from_affine bn254_g2_generator bn254_g2_ecadd
bn254_g2_ecsub bn254_g2_mul_scalar bn254_g2_generate_projective_points bn254_g2_generate_affine_points
bn254_g2_base_field_from_u32
bn254_pairing bn254_pairing_target_field_generate_scalars bn254_pairing_target_field_add bn254_pairing_target_field_sub
bn254_pairing_target_field_mul bn254_pairing_target_field_inv bn254_pairing_target_field_pow
bn254_pairing_target_field_from_u32
""".split()

# the fused fast path: only libicicle_b200 has these
B200_SYMBOLS = """
b200_zkey_cache_create b200_zkey_cache_destroy b200_zkey_cache_info b200_groth16_prove
b200_zkey_cache_create_sharded b200_groth16_commit_partials b200_groth16_finish b200_groth16_prove_files
b200_groth16_commit_begin b200_groth16_commit_end b200_zkey_cache_h_range b200_proof_to_json b200_zkey_cache_b_points
b200_version b200_launch_count b200_fixed_base_mul b200_profile_accumulate b200_profile_records
b200_groth16_verify b200_groth16_verify_files b200_msm_plan_info b200_shard_range
b200_shard_plan b200_shard_plan_mode b200_zkey_cache_ranges b200_comm_unique_id b200_comm_create b200_comm_destroy b200_comm_info b200_groth16_prove_sharded b200_msm_sharded
""".split()


def _ptr(x):
    """numpy array -> void*, int -> void* (device pointer), None -> NULL"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "operands must be contiguous"
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.cast(x, C.c_void_p)


def words(n, w):
    return np.zeros((n, w), dtype=np.uint32)


class IcicleLib:
    """The ICICLE C ABI for BN254, behind whichever shared library `path` names."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)  # RTLD_LOCAL: the reference library exports the same names
        for name in ABI_SYMBOLS:
            fn = getattr(self.dll, name, None)
            if fn is not None and not name.startswith(("bn254_eq", "bn254_is_on", "bn254_g2_eq", "bn254_g2_is_on")):
                fn.restype = C.c_int
        for name in ("bn254_eq", "bn254_is_on_curve", "bn254_g2_eq", "bn254_g2_is_on_curve"):
            getattr(self.dll, name).restype = C.c_bool

    def has(self, name):
        return hasattr(self.dll, name)

    # ---- runtime (rust icicle-runtime: set_device, DeviceVec, IcicleStream) -------------------
    def load_backend_from_env_or_default(self):
        return self.dll.icicle_load_backend_from_env_or_default()

    def set_device(self, type_="CUDA", id_=0):
        d = Device.new(type_, id_)
        check(self.dll.icicle_set_device(C.byref(d)), f"icicle_set_device({type_},{id_})")

    def device_count(self):
        n = C.c_int(0)
        check(self.dll.icicle_get_device_count(C.byref(n)), "icicle_get_device_count")
        return n.value

    def malloc(self, nbytes):
        p = C.c_void_p()
        check(self.dll.icicle_malloc(C.byref(p), C.c_size_t(nbytes)), "icicle_malloc")
        return p.value

    def free(self, ptr):
        check(self.dll.icicle_free(C.c_void_p(ptr)), "icicle_free")

    def copy_to_device(self, dptr, arr):
        check(self.dll.icicle_copy_to_device(C.c_void_p(dptr), _ptr(arr), C.c_size_t(arr.nbytes)), "copy_to_device")

    def copy_to_host(self, arr, dptr):
        check(self.dll.icicle_copy_to_host(_ptr(arr), C.c_void_p(dptr), C.c_size_t(arr.nbytes)), "copy_to_host")

    def is_active_device_memory(self, ptr):
        return self.dll.icicle_is_active_device_memory(C.c_void_p(ptr)) == 0

    def is_host_memory(self, ptr):
        return self.dll.icicle_is_host_memory(C.c_void_p(ptr)) == 0

    def create_stream(self):
        s = C.c_void_p()
        check(self.dll.icicle_create_stream(C.byref(s)), "icicle_create_stream")
        return s.value

    def stream_synchronize(self, s):
        check(self.dll.icicle_stream_synchronize(C.c_void_p(s)), "icicle_stream_synchronize")

    def destroy_stream(self, s):
        check(self.dll.icicle_destroy_stream(C.c_void_p(s)), "icicle_destroy_stream")

    def device_synchronize(self):
        check(self.dll.icicle_device_synchronize(), "icicle_device_synchronize")

    # ---- msm (rust icicle-core msm::msm) ------------------------------------------------------
    def msm(self, scalars, points, cfg=None, g2=False, results=None, msm_size=None):
        """scalars (batch*n, 8) u32; points (n*f[*batch], 16|32) u32; returns (batch, 24|48) u32."""
        cfg = cfg or MSMConfig.default()
        batch = max(1, cfg.batch_size)
        if msm_size is None:
            msm_size = scalars.shape[0] // batch
            # argument checks of msm/mod.rs:106-154
            npts = points.shape[0]
            if msm_size and npts % (msm_size * max(1, cfg.precompute_factor)) != 0 and npts % msm_size != 0:
                raise ValueError("number of points is not a multiple of the msm size")
        pw = G2_PROJ_WORDS if g2 else G1_PROJ_WORDS
        out = results if results is not None else words(batch, pw)
        fn = self.dll.bn254_g2_msm if g2 else self.dll.bn254_msm
        check(fn(_ptr(scalars), _ptr(points), C.c_int(msm_size), C.byref(cfg), _ptr(out)), "bn254_msm")
        return out

    def msm_precompute_bases(self, points, cfg, g2=False, n=None, out=None):
        sets = 1 if (cfg.are_points_shared_in_batch or cfg.batch_size < 1) else cfg.batch_size
        n = points.shape[0] // sets if n is None else n  # per-MSM size; the table covers n * sets points (msm/mod.rs:156-197)
        aw = G2_AFFINE_WORDS if g2 else G1_AFFINE_WORDS
        out = words(n * sets * cfg.precompute_factor, aw) if out is None else out
        fn = self.dll.bn254_g2_msm_precompute_bases if g2 else self.dll.bn254_msm_precompute_bases
        check(fn(_ptr(points), C.c_int(n), C.byref(cfg), _ptr(out)), "bn254_msm_precompute_bases")
        return out

    # ---- ntt (rust icicle-core ntt::{initialize_domain, ntt, ntt_inplace, get_root_of_unity}) --
    def get_root_of_unity(self, max_size):
        r = words(1, 8)
        check(self.dll.bn254_get_root_of_unity(C.c_uint64(max_size), _ptr(r)), "bn254_get_root_of_unity")
        return r[0]

    def ntt_init_domain(self, root):
        cfg = NTTInitDomainConfig(None, False, None)
        root = np.ascontiguousarray(root, dtype=np.uint32)
        check(self.dll.bn254_ntt_init_domain(_ptr(root), C.byref(cfg)), "bn254_ntt_init_domain")

    def ntt_release_domain(self):
        check(self.dll.bn254_ntt_release_domain(), "bn254_ntt_release_domain")

    def ntt(self, inp, direction, cfg=None, out=None, size=None):
        cfg = cfg or NTTConfig.default()
        batch = max(1, cfg.batch_size)
        if size is None:
            size = inp.shape[0] // batch
        out = np.empty_like(inp) if out is None else out
        check(self.dll.bn254_ntt(_ptr(inp), C.c_int(size), C.c_int(direction), C.byref(cfg), _ptr(out)), "bn254_ntt")
        return out

    # ---- vec ops (rust icicle-core vec_ops::{mul_scalars, sub_scalars, ...}) --------------------
    def _vv(self, name, a, b, cfg=None, out=None, n=None):
        cfg = cfg or VecOpsConfig.default()
        batch = max(1, cfg.batch_size)
        n = a.shape[0] // batch if n is None else n
        out = np.empty_like(a) if out is None else out
        check(getattr(self.dll, name)(_ptr(a), _ptr(b), C.c_uint64(n), C.byref(cfg), _ptr(out)), name)
        return out

    def vector_add(self, a, b, **kw):
        return self._vv("bn254_vector_add", a, b, **kw)

    def vector_sub(self, a, b, **kw):
        return self._vv("bn254_vector_sub", a, b, **kw)

    def vector_mul(self, a, b, **kw):
        return self._vv("bn254_vector_mul", a, b, **kw)

    def vector_div(self, a, b, **kw):
        return self._vv("bn254_vector_div", a, b, **kw)

    def vector_accumulate(self, a, b, cfg=None):
        cfg = cfg or VecOpsConfig.default()
        n = a.shape[0] // max(1, cfg.batch_size)
        check(self.dll.bn254_vector_accumulate(_ptr(a), _ptr(b), C.c_uint64(n), C.byref(cfg)), "vector_accumulate")
        return a

    def _reduce(self, name, a, cfg=None):
        cfg = cfg or VecOpsConfig.default()
        batch = max(1, cfg.batch_size)
        out = words(batch, 8)
        check(getattr(self.dll, name)(_ptr(a), C.c_uint64(a.shape[0] // batch), C.byref(cfg), _ptr(out)), name)
        return out

    def vector_sum(self, a, cfg=None):
        return self._reduce("bn254_vector_sum", a, cfg)

    def vector_product(self, a, cfg=None):
        return self._reduce("bn254_vector_product", a, cfg)

    def _sv(self, name, s, v, cfg=None):
        cfg = cfg or VecOpsConfig.default()
        n = v.shape[0] // max(1, cfg.batch_size)
        out = np.empty_like(v)
        check(getattr(self.dll, name)(_ptr(s), _ptr(v), C.c_uint64(n), C.byref(cfg), _ptr(out)), name)
        return out

    def scalar_add_vec(self, s, v, cfg=None):
        return self._sv("bn254_scalar_add_vec", s, v, cfg)

    def scalar_sub_vec(self, s, v, cfg=None):
        return self._sv("bn254_scalar_sub_vec", s, v, cfg)

    def scalar_mul_vec(self, s, v, cfg=None):
        return self._sv("bn254_scalar_mul_vec", s, v, cfg)

    def convert_montgomery(self, x, is_into, kind="scalar", cfg=None, out=None, n=None):
        """kind: scalar | affine | projective | g2_affine | g2_projective (MontgomeryConvertible)"""
        cfg = cfg or VecOpsConfig.default()
        name = {"scalar": "bn254_scalar_convert_montgomery", "affine": "bn254_affine_convert_montgomery",
                "projective": "bn254_projective_convert_montgomery",
                "g2_affine": "bn254_g2_affine_convert_montgomery",
                "g2_projective": "bn254_g2_projective_convert_montgomery"}[kind]
        n = x.shape[0] if n is None else n
        out = np.empty_like(x) if out is None else out
        fn = getattr(self.dll, name)
        nn = C.c_uint64(n) if kind == "scalar" else C.c_size_t(n)
        check(fn(_ptr(x), nn, C.c_bool(is_into), C.byref(cfg), _ptr(out)), name)
        return out

    # ---- host scalar helpers (rust Field/Curve traits over FFI) ---------------------------------
    def _f3(self, name, a, b):
        o = np.zeros(8, dtype=np.uint32)
        getattr(self.dll, name)(_ptr(np.ascontiguousarray(a, dtype=np.uint32)),
                                _ptr(np.ascontiguousarray(b, dtype=np.uint32)), _ptr(o))
        return o

    def fr_add(self, a, b):
        return self._f3("bn254_add", a, b)

    def fr_sub(self, a, b):
        return self._f3("bn254_sub", a, b)

    def fr_mul(self, a, b):
        return self._f3("bn254_mul", a, b)

    def fr_inv(self, a):
        o = np.zeros(8, dtype=np.uint32)
        self.dll.bn254_inv(_ptr(np.ascontiguousarray(a, dtype=np.uint32)), _ptr(o))
        return o

    def fr_pow(self, a, e):
        o = np.zeros(8, dtype=np.uint32)
        self.dll.bn254_pow(_ptr(np.ascontiguousarray(a, dtype=np.uint32)), C.c_int(e), _ptr(o))
        return o

    def generate_scalars(self, n):
        o = words(n, 8)
        self.dll.bn254_generate_scalars(_ptr(o), C.c_int(n))
        return o

    def generate_affine_points(self, n, g2=False):
        o = words(n, G2_AFFINE_WORDS if g2 else G1_AFFINE_WORDS)
        (self.dll.bn254_g2_generate_affine_points if g2 else self.dll.bn254_generate_affine_points)(_ptr(o), C.c_int(n))
        return o

    def _pfx(self, g2):
        return "bn254_g2_" if g2 else "bn254_"

    def generator(self, g2=False):
        o = np.zeros(G2_PROJ_WORDS if g2 else G1_PROJ_WORDS, dtype=np.uint32)
        getattr(self.dll, self._pfx(g2) + "generator")(_ptr(o))
        return o

    def eq(self, a, b, g2=False):
        return bool(getattr(self.dll, self._pfx(g2) + "eq")(_ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b))))

    def is_on_curve(self, p, g2=False):
        return bool(getattr(self.dll, self._pfx(g2) + "is_on_curve")(_ptr(np.ascontiguousarray(p))))

    def to_affine(self, p, g2=False):
        o = np.zeros(G2_AFFINE_WORDS if g2 else G1_AFFINE_WORDS, dtype=np.uint32)
        getattr(self.dll, self._pfx(g2) + "to_affine")(_ptr(np.ascontiguousarray(p)), _ptr(o))
        return o

    def from_affine(self, a, g2=False):
        o = np.zeros(G2_PROJ_WORDS if g2 else G1_PROJ_WORDS, dtype=np.uint32)
        getattr(self.dll, self._pfx(g2) + "from_affine")(_ptr(np.ascontiguousarray(a)), _ptr(o))
        return o

    def ecadd(self, a, b, g2=False):
        o = np.zeros_like(a)
        getattr(self.dll, self._pfx(g2) + "ecadd")(_ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b)), _ptr(o))
        return o

    def ecsub(self, a, b, g2=False):
        o = np.zeros_like(a)
        getattr(self.dll, self._pfx(g2) + "ecsub")(_ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b)), _ptr(o))
        return o

    # ---- pairing (rust icicle-core/src/pairing/mod.rs: pairing(), PairingTargetField) ----------------
    def pairing(self, p_affine, q_affine):
        """e(P, Q) as 96 words (Fq12, standard form); host code in the reference too (icicle/src/pairing.cpp:20-24)"""
        o = np.zeros(FQ12_WORDS, dtype=np.uint32)
        self.dll.bn254_pairing(_ptr(np.ascontiguousarray(p_affine, dtype=np.uint32)),
                               _ptr(np.ascontiguousarray(q_affine, dtype=np.uint32)), _ptr(o))
        return o

    def _t3(self, name, a, b):
        o = np.zeros(FQ12_WORDS, dtype=np.uint32)
        getattr(self.dll, name)(_ptr(np.ascontiguousarray(a, dtype=np.uint32)),
                                _ptr(np.ascontiguousarray(b, dtype=np.uint32)), _ptr(o))
        return o

    def target_add(self, a, b):
        return self._t3("bn254_pairing_target_field_add", a, b)

    def target_sub(self, a, b):
        return self._t3("bn254_pairing_target_field_sub", a, b)

    def target_mul(self, a, b):
        return self._t3("bn254_pairing_target_field_mul", a, b)

    def target_inv(self, a):
        o = np.zeros(FQ12_WORDS, dtype=np.uint32)
        self.dll.bn254_pairing_target_field_inv(_ptr(np.ascontiguousarray(a, dtype=np.uint32)), _ptr(o))
        return o

    def target_pow(self, a, e):
        o = np.zeros(FQ12_WORDS, dtype=np.uint32)
        self.dll.bn254_pairing_target_field_pow(_ptr(np.ascontiguousarray(a, dtype=np.uint32)), C.c_int(e), _ptr(o))
        return o

    def target_from_u32(self, v):
        o = np.zeros(FQ12_WORDS, dtype=np.uint32)
        self.dll.bn254_pairing_target_field_from_u32(C.c_uint32(v), _ptr(o))
        return o

    def target_generate(self, n):
        o = words(n, FQ12_WORDS)
        self.dll.bn254_pairing_target_field_generate_scalars(_ptr(o), C.c_int(n))
        return o

    def mul_scalar(self, p, s, g2=False):
        o = np.zeros_like(p)
        getattr(self.dll, self._pfx(g2) + "mul_scalar")(
            _ptr(np.ascontiguousarray(p)), _ptr(np.ascontiguousarray(s, dtype=np.uint32)), _ptr(o))
        return o
