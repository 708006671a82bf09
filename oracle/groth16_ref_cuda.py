"""The reference's Rust host (src/cache.rs, src/proof_helper.rs) restated call for call over the reference's OWN CUDA
backend (oracle/_ref_cuda, built by oracle/Makefile.ref_cuda for sm_100a and loaded into the reference frontend of
oracle/_ref with icicle_load_backend), with the same residency as the Rust code: zkey points, coefficients and coset
keys live in device memory across proofs (cache.rs:183-231), the witness slices and the scattered A/B rows cross PCIe
every proof (proof_helper.rs:44-104, 194-196).

TEST / BASELINE INFRASTRUCTURE ONLY: the second oracle and `bench.py --impl reference-cuda`, the
"reference CUDA backend on the same B200" baseline.  Never imported by icicle-snark_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import bn254_py as O
from . import groth16_ref as G
from . import ref_cpu

R = O.R_MOD
CUDA_DIR = os.path.join(ref_cpu.ORACLE_DIR, "_ref_cuda")
CUDA_LIB = os.path.join(CUDA_DIR, "libicicle_backend_cuda_device_ref.so")


def available():
    return ref_cpu.available() and os.path.exists(CUDA_LIB)


def build(jobs=8):
    import subprocess
    if os.path.exists(CUDA_LIB):
        return CUDA_LIB
    if not os.path.isdir(ref_cpu.REFERENCE_TREE):
        raise FileNotFoundError(f"{CUDA_LIB} is not built and {ref_cpu.REFERENCE_TREE} is not present on this machine")
    ref_cpu.build_ref(jobs)
    subprocess.check_call(["make", "-C", ref_cpu.ORACLE_DIR, "-f", "Makefile.ref_cuda", f"-j{jobs}"])
    return CUDA_LIB


_loaded = False


def ref_cuda(device_id=0):
    """The reference library object (oracle.ref_cpu.ref()) with its CUDA backend loaded and device set to CUDA."""
    global _loaded
    ref = ref_cpu.ref()
    if not _loaded:
        ref.dll.icicle_load_backend.argtypes = [C.c_char_p, C.c_bool]
        rc = ref.dll.icicle_load_backend(CUDA_DIR.encode(), C.c_bool(False))
        if rc != 0:
            raise RuntimeError(f"icicle_load_backend({CUDA_DIR}) failed: {rc}")
        _loaded = True
    ref.set_device("CUDA", device_id)
    return ref


class DevBuf:
    def __init__(self, ref, nbytes):
        self.ref, self.nbytes = ref, nbytes
        self.ptr = ref.malloc(max(nbytes, 16))

    @classmethod
    def of(cls, ref, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(ref, arr.nbytes)
        if arr.nbytes:
            ref.copy_to_device(b.ptr, arr)
        return b

    def free(self):
        if self.ptr:
            self.ref.free(self.ptr)
            self.ptr = 0


def _vcfg(B, a_dev, b_dev, r_dev):
    cfg = B.VecOpsConfig.default()
    cfg.is_a_on_device, cfg.is_b_on_device, cfg.is_result_on_device = a_dev, b_dev, r_dev
    return cfg


class ZKeyCacheCuda:
    """CacheManager::compute (cache.rs:117-241): everything the proofs reuse, device-resident."""

    def __init__(self, ref, B, zkey_bytes):
        self.ref, self.B = ref, B
        z = self.z = G.parse_zkey(zkey_bytes)
        cfg = _vcfg(B, True, True, True)

        def points(arr, kind):
            d = DevBuf.of(ref, arr)
            if len(arr):
                ref.convert_montgomery(d.ptr, False, kind=kind, cfg=cfg, out=d.ptr, n=len(arr))  # from_mont in place (cache.rs:208-213)
            return d

        self.points_a = points(z["A"], "affine")
        self.points_b1 = points(z["B1"], "affine")
        self.points_b = points(z["B2"], "g2_affine")
        self.points_c = points(z["C"], "affine")
        self.points_h = points(z["H"], "affine")
        self.first_slice = DevBuf.of(ref, z["coef"])
        ref.convert_montgomery(self.first_slice.ptr, False, cfg=cfg, out=self.first_slice.ptr, n=len(z["coef"]))
        host = ref_cpu.ref()
        host.set_device("CPU", 0)
        for k in ("alpha1", "beta1", "delta1"):
            setattr(self, k, host.from_affine(host.convert_montgomery(z[k].reshape(1, 16), False, kind="affine")[0]))
        for k in ("beta2", "gamma2", "delta2"):
            setattr(self, k, host.from_affine(host.convert_montgomery(z[k].reshape(1, 32), False, kind="g2_affine")[0], g2=True))
        ref.set_device("CUDA", 0)
        inc = O.omega(z["power"] + 1)
        keys, cur = [], 1
        for _ in range(z["domain_size"]):
            keys.append(cur)
            cur = cur * inc % R
        self.keys = DevBuf.of(ref, np.frombuffer(b"".join(k.to_bytes(32, "little") for k in keys), dtype=np.uint32).reshape(-1, 8))
        self.root = ref.get_root_of_unity(z["domain_size"])
        # property of the zkey, decided once: does the host scatter (proof_helper.rs:81-92) ever add two values?
        self.scatter_idx = z["c"] + z["m"] * z["domain_size"]
        self.scatter_is_assignment = len(np.unique(self.scatter_idx)) == len(self.scatter_idx)
        ref.ntt_release_domain()
        ref.ntt_init_domain(self.root)

    def close(self):
        for k in ("points_a", "points_b1", "points_b", "points_c", "points_h", "first_slice", "keys"):
            getattr(self, k).free()


def construct_r1cs(ref, B, cache: ZKeyCacheCuda, witness: np.ndarray):
    """proof_helper.rs:31-170; returns the device buffer d_vec (3N scalars)."""
    z = cache.z
    N = z["domain_size"]
    n_coef = len(z["s"])
    second = np.ascontiguousarray(witness[z["s"]])                       # host gather (:52-60)
    d_second = DevBuf.of(ref, second)                                    # copy_from_host_async (:72)
    dd = _vcfg(B, True, True, True)
    ref.convert_montgomery(d_second.ptr, False, cfg=dd, out=d_second.ptr, n=n_coef)  # from_mont (:74)
    res = np.empty((n_coef, 8), dtype=np.uint32)
    ref._vv("bn254_vector_mul", cache.first_slice.ptr, d_second.ptr, cfg=_vcfg(B, True, True, False), out=res, n=n_coef)  # (:75) result on host
    d_second.free()
    idx = cache.scatter_idx
    buf = np.zeros((2 * N, 8), dtype=np.uint32)
    if cache.scatter_is_assignment:                                      # the host scatter loop (:81-92)
        buf[idx] = res
    else:
        out = [0] * (2 * N)
        for i in range(n_coef):
            out[int(idx[i])] = (out[int(idx[i])] + int.from_bytes(res[i].tobytes(), "little")) % R
        buf = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in out), dtype=np.uint32).reshape(-1, 8)
    d_vec = DevBuf(ref, 3 * N * 32)
    ref.copy_to_device(d_vec.ptr, np.ascontiguousarray(buf[N:]))         # (:94-99)
    ref.copy_to_device(d_vec.ptr + N * 32, np.ascontiguousarray(buf[:N]))
    p0, p1, p2 = d_vec.ptr, d_vec.ptr + N * 32, d_vec.ptr + 2 * N * 32
    ref._vv("bn254_vector_mul", p0, p1, cfg=dd, out=p2, n=N)             # (:108-114)
    ncfg = B.NTTConfig.default()
    ncfg.batch_size = 3
    ncfg.are_inputs_on_device = ncfg.are_outputs_on_device = True
    ref.ntt(d_vec.ptr, B.kInverse, ncfg, out=d_vec.ptr, size=N)          # ntt_helper(inverse) (:116)
    for p in (p0, p1, p2):                                               # coset keys (:118-143)
        ref._vv("bn254_vector_mul", p, cache.keys.ptr, cfg=dd, out=p, n=N)
    ref.ntt(d_vec.ptr, B.kForward, ncfg, out=d_vec.ptr, size=N)          # (:145-148)
    ref._vv("bn254_vector_mul", p0, p1, cfg=dd, out=p0, n=N)             # L * R (:154-160)
    ref._vv("bn254_vector_sub", p0, p2, cfg=dd, out=p1, n=N)             # - O  (:161-167)
    return d_vec


def groth16_commitments(ref, B, cache: ZKeyCacheCuda, d_vec: DevBuf, witness: np.ndarray):
    """proof_helper.rs:172-241: witness to the device once, five MSMs on two streams, results copied back."""
    z = cache.z
    N = z["domain_size"]
    d_scalars = DevBuf.of(ref, witness)
    s1, s2 = ref.create_stream(), ref.create_stream()
    outs = []

    def msm(scalars_ptr, n, points: DevBuf, stream, g2=False):
        res = DevBuf(ref, 192 if g2 else 96)
        cfg = B.MSMConfig.default()
        cfg.stream, cfg.is_async = stream, True
        cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
        if n:
            ref.msm(scalars_ptr, points.ptr, cfg, g2=g2, results=res.ptr, msm_size=n)
        outs.append((res, g2, n))

    nw = len(witness)
    msm(d_scalars.ptr, nw, cache.points_a, s1)
    msm(d_scalars.ptr, nw, cache.points_b1, s1)
    nc = nw - (z["n_public"] + 1)
    msm(d_scalars.ptr + (z["n_public"] + 1) * 32, nc, cache.points_c, s1)
    msm(d_vec.ptr + N * 32, N, cache.points_h, s1)
    msm(d_scalars.ptr, nw, cache.points_b, s2, g2=True)
    ref.stream_synchronize(s1)
    ref.stream_synchronize(s2)
    host = []
    for res, g2, n in outs:
        o = np.zeros(48 if g2 else 24, dtype=np.uint32)
        if n:
            ref.copy_to_host(o, res.ptr)
        else:
            o[8] = 1  # identity (0, 1, 0)
        res.free()
        host.append(o)
    ref.destroy_stream(s1)
    ref.destroy_stream(s2)
    d_scalars.free()
    a, b1, c, h, b = host
    return a, b1, b, c, h


def prove(ref, B, wtns_bytes, r: int, s: int, cache: ZKeyCacheCuda, timings: dict | None = None):
    """groth16_prove_helper (proof_helper.rs:243-317) on the CUDA device, epilogue on the host as in the Rust code."""
    z = cache.z
    w = G.parse_wtns(wtns_bytes)
    if w["n_witness"] != z["n_vars"]:
        raise ValueError(f"Invalid witness length. Circuit: {z['n_vars']}, witness: {w['n_witness']}")
    witness = w["w"]
    t0 = time.perf_counter()
    d_vec = construct_r1cs(ref, B, cache, witness)
    ref.device_synchronize()
    t1 = time.perf_counter()
    a, b1, b, c, h = groth16_commitments(ref, B, cache, d_vec, witness)
    d_vec.free()
    t2 = time.perf_counter()
    rw = np.frombuffer((r % R).to_bytes(32, "little"), dtype=np.uint32)
    sw = np.frombuffer((s % R).to_bytes(32, "little"), dtype=np.uint32)
    rsw = np.frombuffer((r * s % R).to_bytes(32, "little"), dtype=np.uint32)
    pi_a = ref.ecadd(ref.ecadd(a, cache.alpha1), ref.mul_scalar(cache.delta1, rw))
    pi_b = ref.ecadd(ref.ecadd(b, cache.beta2, g2=True), ref.mul_scalar(cache.delta2, sw, g2=True), g2=True)
    pi_b1 = ref.ecadd(ref.ecadd(b1, cache.beta1), ref.mul_scalar(cache.delta1, sw))
    pi_c = ref.ecadd(ref.ecadd(ref.ecadd(c, h), ref.mul_scalar(pi_a, sw)), ref.mul_scalar(pi_b1, rw))
    pi_c = ref.ecsub(pi_c, ref.mul_scalar(cache.delta1, rsw))
    if timings is not None:
        timings.update(r1cs_ntt_s=t1 - t0, msm_s=t2 - t1, total_s=time.perf_counter() - t0)
    proof = dict(pi_a=ref.to_affine(pi_a), pi_b=ref.to_affine(pi_b, g2=True), pi_c=ref.to_affine(pi_c))
    public = [int.from_bytes(witness[i].tobytes(), "little") for i in range(1, z["n_public"] + 1)]
    return proof, public
