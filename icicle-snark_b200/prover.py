"""Host-side mirror of the reference's prover API over the fused C ABI (b200_groth16_*).

  groth16_prove(witness, zkey, proof, public, device, cache_manager)   <-> /root/reference/src/lib.rs:33-61
  CacheManager / ZKeyCache                                             <-> /root/reference/src/cache.rs:58-72,110-256

Same argument meaning and error behaviour: paths in, proof.json / public.json out, cache keyed
"{zkey}_{device}", wrong witness length or curve raises (the Rust code panics).  No arithmetic here.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import time

import numpy as np

from .bindings import Groth16Partials, Groth16Proof, IcicleError, ProveTimings, check


def _dec(words):
    return str(int.from_bytes(bytes(words), "little"))


class ZKeyCache:
    """Device-resident, Montgomery-form zkey (points, R1CS in CSR, coset powers, twiddles, workspace)."""

    def __init__(self, lib, zkey_bytes, precompute=1, rank=0, world=1):
        self.lib = lib
        self.handle = C.c_void_p()
        if isinstance(zkey_bytes, bytes):
            buf = C.c_char_p(zkey_bytes)  # the object's own buffer: no copy of a gigabyte-sized file
        elif isinstance(zkey_bytes, mmap.mmap):
            buf = (C.c_ubyte * len(zkey_bytes)).from_buffer(zkey_bytes)
        else:
            buf = (C.c_ubyte * len(zkey_bytes)).from_buffer_copy(zkey_bytes)
        check(lib.dll.b200_zkey_cache_create_sharded(buf, C.c_size_t(len(zkey_bytes)), C.c_int(precompute), C.c_int(rank),
                                                     C.c_int(world), C.byref(self.handle)), "b200_zkey_cache_create")
        nv, npub, dom, ncoef, dbytes = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        check(lib.dll.b200_zkey_cache_info(self.handle, C.byref(nv), C.byref(npub), C.byref(dom), C.byref(ncoef),
                                           C.byref(dbytes)))
        self.n_vars, self.n_public, self.domain_size = nv.value, npub.value, dom.value
        self.n_coef, self.device_bytes = ncoef.value, dbytes.value
        self.rank, self.world = rank, world

    def close(self):
        if self.handle:
            self.lib.dll.b200_zkey_cache_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # witness: (n_vars, 8) u32 numpy array (standard form) or a raw host pointer (int) + count
    def prove(self, witness, r=None, s=None, n_witness=None):
        proof, tm = Groth16Proof(), ProveTimings()
        wp, n = self._wptr(witness, n_witness)
        rp = None if r is None else np.frombuffer(int(r).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        sp = None if s is None else np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        check(self.lib.dll.b200_groth16_prove(self.handle, wp, C.c_uint32(n), rp, sp, C.byref(proof), C.byref(tm)),
              "b200_groth16_prove")
        return proof, tm

    def prove_sharded(self, comm, witness, r=None, s=None, n_witness=None):
        """One proof over the communicator's GPUs with the whole exchange inside the library
        (b200_groth16_prove_sharded); returns (proof on rank 0 / None elsewhere, timings)."""
        proof, tm = Groth16Proof(), ProveTimings()
        wp, n = self._wptr(witness, n_witness)
        rp = None if r is None else np.frombuffer(int(r).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        sp = None if s is None else np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        check(self.lib.dll.b200_groth16_prove_sharded(self.handle, comm.handle, wp, C.c_uint32(n), rp, sp, C.byref(proof), C.byref(tm)),
              "b200_groth16_prove_sharded")
        return (proof if comm.rank == 0 else None), tm

    def commit_partials(self, witness, n_witness=None):
        parts, tm = Groth16Partials(), ProveTimings()
        wp, n = self._wptr(witness, n_witness)
        check(self.lib.dll.b200_groth16_commit_partials(self.handle, wp, C.c_uint32(n), C.byref(parts), C.byref(tm)),
              "b200_groth16_commit_partials")
        return parts, tm

    def h_range(self):
        lo, hi = C.c_uint32(), C.c_uint32()
        check(self.lib.dll.b200_zkey_cache_h_range(self.handle, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def ranges(self):
        """[(lo, hi)] x 5: the part of the sections (H, A, B1, C, B2) this cache holds (b200_zkey_cache_ranges)."""
        lo, hi = (C.c_uint32 * 5)(), (C.c_uint32 * 5)()
        check(self.lib.dll.b200_zkey_cache_ranges(self.handle, lo, hi))
        return [(int(lo[k]), int(hi[k])) for k in range(5)]

    def b_points(self):
        kept, total = C.c_uint32(), C.c_uint32()
        check(self.lib.dll.b200_zkey_cache_b_points(self.handle, C.byref(kept), C.byref(total)))
        return kept.value, total.value

    def commit_begin(self, witness, first_poly, poly_count, out_dev_ptr, n_witness=None):
        """N > 1 quotient split, phase 1 (see include/icicle_b200.h)."""
        wp, n = self._wptr(witness, n_witness)
        check(self.lib.dll.b200_groth16_commit_begin(self.handle, wp, C.c_uint32(n), C.c_int(first_poly), C.c_int(poly_count),
                                                     C.c_void_p(out_dev_ptr)), "b200_groth16_commit_begin")

    def commit_end(self, a_dev_ptr, b_dev_ptr, c_dev_ptr):
        parts, tm = Groth16Partials(), ProveTimings()
        check(self.lib.dll.b200_groth16_commit_end(self.handle, C.c_void_p(a_dev_ptr), C.c_void_p(b_dev_ptr), C.c_void_p(c_dev_ptr),
                                                   C.byref(parts), C.byref(tm)), "b200_groth16_commit_end")
        return parts, tm

    def finish(self, parts_list, r=None, s=None):
        arr = (Groth16Partials * len(parts_list))(*parts_list)
        proof = Groth16Proof()
        rp = None if r is None else np.frombuffer(int(r).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        sp = None if s is None else np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint32).ctypes.data_as(C.c_void_p)
        check(self.lib.dll.b200_groth16_finish(self.handle, arr, C.c_int(len(parts_list)), rp, sp, C.byref(proof)),
              "b200_groth16_finish")
        return proof

    @staticmethod
    def _wptr(witness, n_witness):
        if isinstance(witness, np.ndarray):
            assert witness.flags["C_CONTIGUOUS"] and witness.dtype == np.uint32
            return witness.ctypes.data_as(C.c_void_p), witness.shape[0] if n_witness is None else n_witness
        return C.c_void_p(int(witness)), n_witness


def proof_to_dict(proof: Groth16Proof):
    a, b, c = list(proof.pi_a), list(proof.pi_b), list(proof.pi_c)
    w = lambda x: np.array(x, dtype=np.uint32)
    return dict(pi_a=w(a), pi_b=w(b), pi_c=w(c))


def proof_json(proof: Groth16Proof) -> str:
    """Same bytes the C writer (b200_groth16_prove_files) and serde_json's pretty printer produce."""
    d = proof_to_dict(proof)
    g = lambda p, i: str(int.from_bytes(p[8 * i:8 * i + 8].tobytes(), "little"))
    a, b, c = d["pi_a"], d["pi_b"], d["pi_c"]
    g1 = lambda p: f'[\n    "{g(p, 0)}",\n    "{g(p, 1)}",\n    "1"\n  ]'
    pb = ('[\n    [\n      "%s",\n      "%s"\n    ],\n    [\n      "%s",\n      "%s"\n    ],\n    [\n      "1",\n      "0"\n    ]\n  ]'
          % (g(b, 0), g(b, 1), g(b, 2), g(b, 3)))
    return '{\n  "curve": "bn128",\n  "pi_a": %s,\n  "pi_b": %s,\n  "pi_c": %s,\n  "protocol": "groth16"\n}' % (g1(a), pb, g1(c))


class CacheManager:
    """cache.rs:110-114: one ZKeyCache per "{zkey}_{device}", alive for the process."""

    def __init__(self, lib=None, precompute=1):
        self.lib = lib
        self.precompute = precompute
        self.cache = {}
        self.last_key = ""

    def contains(self, key):
        return key in self.cache

    def compute(self, zkey_path):
        with open(zkey_path, "rb") as f:
            data = f.read()
        return ZKeyCache(self.lib, data, self.precompute)

    def insert_cache(self, key, cache):
        self.cache[key] = cache

    def get_cache(self, key):
        self.last_key = key
        return self.cache[key]


def groth16_prove(witness, zkey, proof, public, device, cache_manager: CacheManager, r=None, s=None):
    """src/lib.rs:33-61.  `device` must be "CUDA": there is no CPU backend behind this library.
    r, s: optional fixed blinding factors (1, 1 == the reference's `no-randomness` feature)."""
    from . import lib as _lib
    lib = cache_manager.lib or _lib()
    cache_manager.lib = lib
    start = time.perf_counter()
    lib.set_device(device, 0)  # raises IcicleError(INVALID_DEVICE) for "CPU"
    key = f"{zkey}_{device}"
    if not cache_manager.contains(key):
        cache_manager.insert_cache(key, cache_manager.compute(zkey))
    cache = cache_manager.get_cache(key)
    with open(witness, "rb") as f:
        wt = f.read()
    # wtns: header section 1 = n8, prime, n_witness; section 2 = witness (file_wrapper.rs:169-177)
    import struct
    if wt[:4] != b"wtns":
        raise ValueError(f"{witness}: Invalid File format")
    nsec = struct.unpack_from("<I", wt, 8)[0]
    pos, secs = 12, {}
    for _ in range(nsec):
        sid, size = struct.unpack_from("<IQ", wt, pos)
        pos += 12
        secs.setdefault(sid, (pos, size))
        pos += size
    p1, _ = secs[1]
    n8 = struct.unpack_from("<I", wt, p1)[0]
    prime = int.from_bytes(wt[p1 + 4:p1 + 4 + n8], "little")
    n_witness = struct.unpack_from("<I", wt, p1 + 4 + n8)[0]
    if prime != 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001:
        raise ValueError("Curve of the witness does not match the curve of the proving key")
    if n_witness != cache.n_vars:
        raise ValueError(f"Invalid witness length. Circuit: {cache.n_vars}, witness: {n_witness}")
    p2, s2 = secs[2]
    w = np.frombuffer(wt, dtype=np.uint32, count=n_witness * 8, offset=p2).reshape(-1, 8)
    pr, tm = cache.prove(np.ascontiguousarray(w), r, s)
    with open(proof, "w") as f:
        f.write(proof_json(pr))
    pub = [int.from_bytes(w[i].tobytes(), "little") for i in range(1, cache.n_public + 1)]
    with open(public, "w") as f:
        f.write("[]" if not pub else "[\n" + ",\n".join(f'  "{x}"' for x in pub) + "\n]")
    print(f"proof took: {time.perf_counter() - start:.6f}s")
    return pr, tm


def groth16_verify(proof, public, vk, lib=None):
    """src/lib.rs:63-82: read proof.json, public.json, verification_key.json; four pairings; raise where the Rust code's
    `assert!(pairing_result, "Verification failed")` panics.  Host code, as in the reference (pairing is CPU-only there)."""
    from . import lib as _lib
    lib = lib or _lib()
    ok = C.c_int(0)
    check(lib.dll.b200_groth16_verify_files(os.fsencode(proof), os.fsencode(public), os.fsencode(vk), C.byref(ok)),
          "b200_groth16_verify_files")
    if not ok.value:
        raise AssertionError("Verification failed")


def groth16_verify_points(lib, proof, public, vk):
    """`groth16_verify_helper` (src/proof_helper.rs:319-372) on in-memory values: proof = Groth16Proof or a dict of
    standard-form pi_a/pi_b/pi_c words, public = ints, vk = dict(alpha1, beta2, gamma2, delta2, ic).  Returns bool."""
    if not isinstance(proof, Groth16Proof):
        p = Groth16Proof()
        for name in ("pi_a", "pi_b", "pi_c"):
            arr = np.ascontiguousarray(proof[name], dtype=np.uint32).reshape(-1)
            C.memmove(getattr(p, name), arr.ctypes.data, arr.nbytes)
        proof = p
    n = len(public)
    pubs = np.zeros((max(n, 1), 8), dtype=np.uint32)
    for i, x in enumerate(public):
        pubs[i] = np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint32)
    ic = np.ascontiguousarray(vk["ic"], dtype=np.uint32)
    if ic.shape[0] < n + 1:
        raise ValueError(f"verification key holds {ic.shape[0]} IC points, {n + 1} needed")
    vp = lambda a: np.ascontiguousarray(a, dtype=np.uint32).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(vk[k], dtype=np.uint32) for k in ("alpha1", "beta2", "gamma2", "delta2")]
    ok = C.c_int(0)
    check(lib.dll.b200_groth16_verify(C.byref(proof), vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), vp(ic), vp(pubs),
                                      C.c_uint64(n), C.byref(ok)), "b200_groth16_verify")
    return bool(ok.value)
