"""CPU restatement of the reference's Rust host for the Groth16 path, driving the reference's own C++
(oracle/_ref, via the shared C ABI) for all arithmetic.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs). The Rust host cannot be
built here (no cargo/rustc), so its glue is restated call for call; the arithmetic underneath is the
reference's compiled code, not a port:

  parse_binfile / parse_zkey / parse_wtns   /root/reference/src/file_wrapper.rs:45-103,169-237, src/zkey.rs:47-85
  ZKeyCacheRef (compute)                    /root/reference/src/cache.rs:117-241 (+ pre_compute_keys 264-289)
  construct_r1cs                            /root/reference/src/proof_helper.rs:31-170
  groth16_commitments                       /root/reference/src/proof_helper.rs:172-241
  prove (epilogue, r/s)                     /root/reference/src/proof_helper.rs:243-317
  proof_json / public_json                  /root/reference/src/conversions.rs:30-56, serde_json pretty
  verify                                    /root/reference/src/proof_helper.rs:319-372 (bn254_pairing)

Parity status: the bn254_* layer is pinned by the compiled reference itself; this glue is pinned by the
pairing check (a wrong call sequence does not verify) - the reference ships no proof fixtures.
"""
from __future__ import annotations

import ctypes as C
import struct
import time

import numpy as np

from . import bn254_py as O

R = O.R_MOD
FR_BYTES = R.to_bytes(32, "little")
FQ_BYTES = O.Q_MOD.to_bytes(32, "little")


# ------------------------------------------------------------------------------------------- formats
def parse_binfile(buf: bytes, magic: bytes, max_version: int = 2):
    if buf[:4] != magic:
        raise ValueError("Invalid File format")
    version, nsec = struct.unpack_from("<II", buf, 4)
    if version > max_version:
        raise ValueError("Version not supported")
    pos, sections = 12, {}
    for _ in range(nsec):
        sid, size = struct.unpack_from("<IQ", buf, pos)
        pos += 12
        sections.setdefault(sid, (pos, size))
        pos += size
    return sections


def _sec(buf, sections, sid):
    p, n = sections[sid]
    return memoryview(buf)[p:p + n]


def words(mv, w):
    return np.frombuffer(mv, dtype=np.uint32).reshape(-1, w).copy()


def parse_zkey(buf: bytes):
    s = parse_binfile(buf, b"zkey")
    if struct.unpack_from("<I", _sec(buf, s, 1), 0)[0] != 1:
        raise ValueError("Protocol not supported")
    h = bytes(_sec(buf, s, 2))
    n8q = struct.unpack_from("<I", h, 0)[0]
    assert n8q == 32 and h[4:36] == FQ_BYTES
    n8r = struct.unpack_from("<I", h, 36)[0]
    assert n8r == 32 and h[40:72] == FR_BYTES
    n_vars, n_public, domain_size = struct.unpack_from("<III", h, 72)
    p = 84
    z = dict(n_vars=n_vars, n_public=n_public, domain_size=domain_size, power=domain_size.bit_length() - 1)
    for name, size in (("alpha1", 64), ("beta1", 64), ("beta2", 128), ("gamma2", 128), ("delta1", 64), ("delta2", 128)):
        z[name] = np.frombuffer(h[p:p + size], dtype=np.uint32).copy()  # Montgomery form, as stored
        p += size
    coefs = _sec(buf, s, 4)
    n_coef = (len(coefs) - 4) // 44
    rec = np.frombuffer(coefs[4:4 + n_coef * 44], dtype=np.uint8).reshape(n_coef, 44)
    z["m"] = rec[:, 0].astype(np.int64)
    z["c"] = rec[:, 4:8].copy().view(np.uint32).reshape(-1).astype(np.int64)
    z["s"] = rec[:, 8:12].copy().view(np.uint32).reshape(-1).astype(np.int64)
    z["coef"] = rec[:, 12:44].copy().view(np.uint32).reshape(-1, 8)
    z["ic"] = words(_sec(buf, s, 3), 16) if 3 in s else None
    z["A"], z["B1"] = words(_sec(buf, s, 5), 16), words(_sec(buf, s, 6), 16)
    z["B2"] = words(_sec(buf, s, 7), 32)
    z["C"], z["H"] = words(_sec(buf, s, 8), 16), words(_sec(buf, s, 9), 16)
    return z


def parse_wtns(buf: bytes):
    s = parse_binfile(buf, b"wtns")
    h = bytes(_sec(buf, s, 1))
    n8 = struct.unpack_from("<I", h, 0)[0]
    q = h[4:4 + n8]
    n_witness = struct.unpack_from("<I", h, 4 + n8)[0]
    return dict(n8=n8, q=q, n_witness=n_witness, w=words(_sec(buf, s, 2), 8))


# ------------------------------------------------------------------------------------------- cache.rs
class ZKeyCacheRef:
    """CacheManager::compute + get_cache on the reference CPU backend."""

    def __init__(self, ref, zkey_bytes):
        self.ref = ref
        z = self.z = parse_zkey(zkey_bytes)
        # points and coefficients leave Montgomery form once (cache.rs:208-214)
        self.points_a = ref.convert_montgomery(z["A"], False, kind="affine")
        self.points_b1 = ref.convert_montgomery(z["B1"], False, kind="affine")
        self.points_b = ref.convert_montgomery(z["B2"], False, kind="g2_affine")
        self.points_c = ref.convert_montgomery(z["C"], False, kind="affine") if len(z["C"]) else z["C"]
        self.points_h = ref.convert_montgomery(z["H"], False, kind="affine")
        self.first_slice = ref.convert_montgomery(z["coef"], False)
        for k in ("alpha1", "beta1", "delta1"):
            setattr(self, k, ref.from_affine(ref.convert_montgomery(z[k].reshape(1, 16), False, kind="affine")[0]))
        for k in ("beta2", "gamma2", "delta2"):
            setattr(self, k, ref.from_affine(
                ref.convert_montgomery(z[k].reshape(1, 32), False, kind="g2_affine")[0], g2=True))
        # keys = [1, inc, inc^2, ...], inc = W[power+1] (cache.rs:168-169, 264-289)
        inc = O.omega(z["power"] + 1)
        keys, cur = [], 1
        for _ in range(z["domain_size"]):
            keys.append(cur)
            cur = cur * inc % R
        self.keys = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in keys), dtype=np.uint32).reshape(-1, 8).copy()
        # get_cache: domain from get_root_of_unity (sized by domain_size here: SURVEY App. C)
        self.root = ref.get_root_of_unity(z["domain_size"])

    def get_cache(self):
        """cache.rs:242-256: (re)initialise the process-wide NTT domain when the zkey changes."""
        if getattr(self.ref, "domain_root", None) != self.root.tobytes():
            self.ref.ntt_release_domain()
            self.ref.ntt_init_domain(self.root)
        return self


def construct_r1cs(ref, cache: ZKeyCacheRef, witness: np.ndarray, B):
    """proof_helper.rs:31-170, same buffers and call order; returns d_vec (3N, 8)."""
    z = cache.z
    N = z["domain_size"]
    second = ref.convert_montgomery(np.ascontiguousarray(witness[z["s"]]), False)  # from_mont on the gathered witness
    res = ref.vector_mul(cache.first_slice, second)
    idx = z["c"] + z["m"] * N
    buf = np.zeros((2 * N, 8), dtype=np.uint32)
    if len(np.unique(idx)) == len(idx):  # no collisions: the scatter loop (:81-92) is a plain assignment
        buf[idx] = res
    else:
        out = [0] * (2 * N)
        res_int = [int.from_bytes(res[i].tobytes(), "little") for i in range(len(res))]
        for i, v in enumerate(res_int):  # the host scatter loop (:81-92), collisions add mod r
            out[int(idx[i])] = (out[int(idx[i])] + v) % R
        buf = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in out), dtype=np.uint32).reshape(-1, 8)
    d = np.zeros((3 * N, 8), dtype=np.uint32)
    d[0:N] = buf[N:]
    d[N:2 * N] = buf[:N]
    d[2 * N:] = ref.vector_mul(np.ascontiguousarray(d[0:N]), np.ascontiguousarray(d[N:2 * N]))
    cfg = B.NTTConfig.default()
    cfg.batch_size = 3
    d = ref.ntt(d, B.kInverse, cfg)
    for k in range(3):
        d[k * N:(k + 1) * N] = ref.vector_mul(np.ascontiguousarray(d[k * N:(k + 1) * N]), cache.keys)
    d = ref.ntt(d, B.kForward, cfg)
    d[0:N] = ref.vector_mul(np.ascontiguousarray(d[0:N]), np.ascontiguousarray(d[N:2 * N]))
    d[N:2 * N] = ref.vector_sub(np.ascontiguousarray(d[0:N]), np.ascontiguousarray(d[2 * N:]))
    return d


def groth16_commitments(ref, cache: ZKeyCacheRef, d_vec, witness):
    """proof_helper.rs:172-241"""
    z = cache.z
    N = z["domain_size"]
    a = ref.msm(witness, cache.points_a)[0]
    b1 = ref.msm(witness, cache.points_b1)[0]
    wc = np.ascontiguousarray(witness[z["n_public"] + 1:])
    c = ref.msm(wc, cache.points_c)[0] if len(wc) else ref.ecsub(a, a)
    h = ref.msm(np.ascontiguousarray(d_vec[N:2 * N]), cache.points_h)[0]
    b = ref.msm(witness, cache.points_b, g2=True)[0]
    return a, b1, b, c, h


def prove(ref, B, zkey_bytes, wtns_bytes, r: int, s: int, cache: ZKeyCacheRef | None = None, timings: dict | None = None):
    """groth16_prove_helper with injected blinding factors (r = s = 1 is the `no-randomness` feature)."""
    cache = cache or ZKeyCacheRef(ref, zkey_bytes)
    z = cache.z
    if hasattr(cache, "get_cache"):
        cache.get_cache()
    w = parse_wtns(wtns_bytes)
    if w["q"] != FR_BYTES:
        raise ValueError("Curve of the witness does not match the curve of the proving key")
    if w["n_witness"] != z["n_vars"]:
        raise ValueError(f"Invalid witness length. Circuit: {z['n_vars']}, witness: {w['n_witness']}")
    witness = w["w"]
    t0 = time.perf_counter()
    d_vec = construct_r1cs(ref, cache, witness, B)
    t1 = time.perf_counter()
    a, b1, b, c, h = groth16_commitments(ref, cache, d_vec, witness)
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


# ------------------------------------------------------------------------------------------- JSON
def _dec(wds):
    return str(int.from_bytes(np.ascontiguousarray(wds, dtype=np.uint32).tobytes(), "little"))


def proof_json(proof) -> str:
    """serde_json::to_writer_pretty(json!(Proof)): alphabetical keys, 2-space indent, no trailing newline."""
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]

    def g1(p):
        return f'[\n    "{_dec(p[:8])}",\n    "{_dec(p[8:16])}",\n    "1"\n  ]'

    pb = ('[\n    [\n      "%s",\n      "%s"\n    ],\n    [\n      "%s",\n      "%s"\n    ],\n    [\n      "1",\n      "0"\n    ]\n  ]'
          % (_dec(b[0:8]), _dec(b[8:16]), _dec(b[16:24]), _dec(b[24:32])))
    return ('{\n  "curve": "bn128",\n  "pi_a": %s,\n  "pi_b": %s,\n  "pi_c": %s,\n  "protocol": "groth16"\n}'
            % (g1(a), pb, g1(c)))


def public_json(public) -> str:
    if not public:
        return "[]"
    return "[\n" + ",\n".join(f'  "{p}"' for p in public) + "\n]"


# ------------------------------------------------------------------------------------------- verify
def _pairing(ref, p_affine, q_affine):
    out = np.zeros(96, dtype=np.uint32)  # Fq12
    ref.dll.bn254_pairing(np.ascontiguousarray(p_affine).ctypes.data_as(C.c_void_p),
                          np.ascontiguousarray(q_affine).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def verify(ref, proof, public, vk) -> bool:
    """e(-A,B) * e(cpub,gamma2) * e(C,delta2) * e(alpha1,beta2) == 1 (proof_helper.rs:338-369).
    vk: dict with standard-form affine alpha1, beta2, gamma2, delta2 and ic (n_public+1, 16)."""
    cpub = ref.from_affine(vk["ic"][0])
    for i, p in enumerate(public):
        sw = np.frombuffer((int(p) % R).to_bytes(32, "little"), dtype=np.uint32)
        cpub = ref.ecadd(cpub, ref.mul_scalar(ref.from_affine(vk["ic"][i + 1]), sw))
    pa = ref.from_affine(proof["pi_a"])
    neg_a = ref.to_affine(ref.ecsub(ref.ecsub(pa, pa), pa))
    fs = [_pairing(ref, neg_a, proof["pi_b"]), _pairing(ref, ref.to_affine(cpub), vk["gamma2"]),
          _pairing(ref, proof["pi_c"], vk["delta2"]), _pairing(ref, vk["alpha1"], vk["beta2"])]
    acc = fs[0]
    for f in fs[1:]:
        nxt = np.zeros(96, dtype=np.uint32)
        ref.dll.bn254_pairing_target_field_mul(acc.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                                               nxt.ctypes.data_as(C.c_void_p))
        acc = nxt
    one = np.zeros(96, dtype=np.uint32)
    ref.dll.bn254_pairing_target_field_from_u32(C.c_uint32(1), one.ctypes.data_as(C.c_void_p))
    return bool(np.array_equal(acc, one))
