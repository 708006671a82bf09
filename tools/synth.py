"""Synthetic Groth16 artefacts for the reference's benchmark circuit, without circom/snarkjs (neither is
available offline): a VALID .zkey / .wtns / verification key for ComplexCircuit(n, n)
(/root/reference/benchmark/*/circuit.circom; input a = 3) from a seeded, known-toxic-waste setup
(SURVEY 8d).  Test/bench infrastructure, not part of the product path.

Signals  w = [1, c, a, b_0 .. b_{n-2}],  b_0 = a^2, b_i = b_{i-1}^2, c = b_{n-1}   (n_vars = n + 2, n_public = 1)
R1CS     row i: A = B = {x_i: 1}, C = {y_i: 1} with x_0 = a, x_i = b_{i-1}, y_i = b_i (y_{n-1} = c);
         rows n, n+1: the snarkjs public-input rows A = {signal 0: 1}, {signal 1: 1}
Setup    A_s = [u_s(tau)]_1, B_s = [v_s(tau)]_{1,2}, C_s = [(beta u_s + alpha v_s + w_s)(tau)/delta]_1 (private s),
         IC_s likewise /gamma, H_j = [L^(2N)_{2j+1}(tau)/delta]_1; Lagrange values by an inverse NTT of the
         tau powers: L_i(tau) = iNTT([tau^k])_i.
File layout: iden3 binfile exactly as the reference reads it (src/file_wrapper.rs:45-103, src/zkey.rs:47-85,
src/cache.rs:126-181): points Montgomery affine, coefficients coef*R^2, witness standard form.

All heavy arithmetic goes through an ICICLE-ABI backend object (bindings.IcicleLib): the product library on
a GPU (plus its b200_fixed_base_mul tool) or the reference CPU library for tiny instances.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import struct

import numpy as np

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
MONT_R = 1 << 256
kForward, kInverse = 0, 1


def ints_to_words(xs):
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in xs), dtype=np.uint32).reshape(-1, 8).copy()


def words_to_ints(a):
    raw = np.ascontiguousarray(a, dtype=np.uint32).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def word(x):
    return ints_to_words([x])


def toxic(seed=b"icicle-snark-b200"):
    vals = [int.from_bytes(hashlib.sha256(seed + bytes([i])).digest(), "big") % R for i in range(5)]
    return dict(zip(("tau", "alpha", "beta", "gamma", "delta"), vals))


def fixed_base(backend, scalars, g2=False):
    """(n,8) standard-form scalars -> (n,16|32) Montgomery-form affine points k*G."""
    n = scalars.shape[0]
    wpp = 32 if g2 else 16
    out = np.zeros((n, wpp), dtype=np.uint32)
    if n == 0:
        return out
    if backend.has("b200_fixed_base_mul"):
        scalars = np.ascontiguousarray(scalars)
        rc = backend.dll.b200_fixed_base_mul(scalars.ctypes.data_as(C.c_void_p), C.c_uint64(n), C.c_int(int(g2)),
                                             C.c_int(1), out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"b200_fixed_base_mul failed: {rc}")
        return out
    gen = backend.generator(g2=g2)
    for i in range(n):  # reference CPU library: tiny instances only
        out[i] = backend.to_affine(backend.mul_scalar(gen, scalars[i], g2=g2), g2=g2)
    return backend.convert_montgomery(out, True, kind="g2_affine" if g2 else "affine")


def lagrange_at_tau(backend, tau, logn):
    """[L_i(tau)]_{i<2^logn} = iNTT of the tau powers; the backend's NTT domain must cover 2^logn."""
    n = 1 << logn
    pw, cur = [], 1
    for _ in range(n):
        pw.append(cur)
        cur = cur * tau % R
    return backend.ntt(ints_to_words(pw), kInverse)


def witness_complex_circuit(n, a=3):
    w = [1, 0, a % R]
    b = a * a % R
    for _ in range(n - 1):
        w.append(b)
        b = b * b % R
    w[1] = b  # c = b_{n-1}
    return w


def _section(sid, payload):
    return struct.pack("<IQ", sid, len(payload)) + payload


def make_complex_circuit(backend, n, seed=b"icicle-snark-b200", a=3, log=None):
    """Returns (zkey_bytes, wtns_bytes, vk) for ComplexCircuit(n, n)."""
    say = log or (lambda *_: None)
    n_public, n_vars = 1, n + 2
    power = (n + n_public + 1 - 1).bit_length()  # smallest 2^power >= n + n_public + 1
    N = 1 << power
    tw = toxic(seed)
    tau, alpha, beta, gamma, delta = (tw[k] for k in ("tau", "alpha", "beta", "gamma", "delta"))
    dinv, ginv = pow(delta, -1, R), pow(gamma, -1, R)

    backend.ntt_release_domain()  # init is a no-op when a (possibly smaller) domain already exists
    backend.ntt_init_domain(backend.get_root_of_unity(2 * N))
    try:
        say("lagrange N")
        L = lagrange_at_tau(backend, tau, power)           # (N, 8)
        say("lagrange 2N")
        L2 = lagrange_at_tau(backend, tau, power + 1)      # (2N, 8)
    finally:
        backend.ntt_release_domain()

    zero = np.zeros((1, 8), dtype=np.uint32)
    # u_s, v_s, w_s at tau per signal s (see module docstring)
    u = np.concatenate([L[n:n + 1], L[n + 1:n + 2], L[:n]])
    v = np.concatenate([zero, zero, L[:n]])
    wv = np.concatenate([zero, L[n - 1:n], zero, L[:n - 1]])
    assert u.shape[0] == v.shape[0] == wv.shape[0] == n_vars
    say("combine")
    comb = backend.vector_add(backend.vector_add(backend.scalar_mul_vec(word(beta), u), backend.scalar_mul_vec(word(alpha), v)), wv)
    c_priv = backend.scalar_mul_vec(word(dinv), np.ascontiguousarray(comb[n_public + 1:]))
    ic_sc = backend.scalar_mul_vec(word(ginv), np.ascontiguousarray(comb[:n_public + 1]))
    h_sc = backend.scalar_mul_vec(word(dinv), np.ascontiguousarray(L2[1::2]))

    say("points A")
    pA = fixed_base(backend, u)
    say("points B1")
    pB1 = fixed_base(backend, v)
    say("points B2")
    pB2 = fixed_base(backend, v, g2=True)
    say("points C")
    pC = fixed_base(backend, c_priv)
    say("points H")
    pH = fixed_base(backend, h_sc)
    pIC = fixed_base(backend, ic_sc)
    vk1 = fixed_base(backend, ints_to_words([alpha, beta, delta]))
    vk2 = fixed_base(backend, ints_to_words([beta, gamma, delta]), g2=True)

    # section 4: coefficient records [m:u32][row:u32][signal:u32][coef*R^2 mod r]
    say("coefficients")
    n_coef = 2 * n + n_public + 1
    rec = np.zeros((n_coef, 44), dtype=np.uint8)
    one_r2 = np.frombuffer((MONT_R * MONT_R % R).to_bytes(32, "little"), dtype=np.uint8)
    rows = np.arange(n, dtype=np.uint32)
    sig = (rows + 2).astype(np.uint32)
    m = np.concatenate([np.zeros(n, np.uint32), np.ones(n, np.uint32), np.zeros(n_public + 1, np.uint32)])
    row = np.concatenate([rows, rows, np.arange(n, n + n_public + 1, dtype=np.uint32)])
    sg = np.concatenate([sig, sig, np.arange(n_public + 1, dtype=np.uint32)])
    rec[:, 0:4] = m.view(np.uint8).reshape(-1, 4)
    rec[:, 4:8] = row.view(np.uint8).reshape(-1, 4)
    rec[:, 8:12] = sg.view(np.uint8).reshape(-1, 4)
    rec[:, 12:44] = one_r2
    sec4 = struct.pack("<I", n_coef) + rec.tobytes()

    header = (struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little") +
              struct.pack("<III", n_vars, n_public, N) + vk1[0].tobytes() + vk1[1].tobytes() + vk2[0].tobytes() +
              vk2[1].tobytes() + vk1[2].tobytes() + vk2[2].tobytes())
    sections = [(1, struct.pack("<I", 1)), (2, header), (3, pIC.tobytes()), (4, sec4), (5, pA.tobytes()), (6, pB1.tobytes()),
                (7, pB2.tobytes()), (8, pC.tobytes()), (9, pH.tobytes())]
    zkey = b"zkey" + struct.pack("<II", 1, len(sections)) + b"".join(_section(i, p) for i, p in sections)

    say("witness")
    w = witness_complex_circuit(n, a)
    wt_hdr = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", n_vars)
    wtns = (b"wtns" + struct.pack("<II", 2, 2) + _section(1, wt_hdr) +
            _section(2, b"".join(x.to_bytes(32, "little") for x in w)))

    std1 = backend.convert_montgomery(vk1, False, kind="affine")
    std2 = backend.convert_montgomery(vk2, False, kind="g2_affine")
    vk = dict(alpha1=std1[0], beta2=std2[0], gamma2=std2[1], delta2=std2[2],
              ic=backend.convert_montgomery(pIC, False, kind="affine"), n_public=n_public)
    return zkey, wtns, vk


def make_random_circuit(backend, n_constraints, n_inputs=64, n_public=9, max_terms=6, zero_one_fraction=0.9, seed=b"aadhaar-shaped",
                        rng_seed=7):
    """"aadhaar-shaped" SUBSTITUTE for benchmark/anon_aadhaar (which needs circom + circomlib + snarkjs, unavailable
    offline): a satisfiable random R1CS with multi-entry rows and a 0/1-heavy witness.
    Constraint k multiplies two random sparse linear combinations of earlier signals (1..max_terms terms, small
    coefficients) and defines a NEW signal as the product, so every constraint is satisfied by construction:
        (sum a_t w_{i_t}) * (sum b_t w_{j_t}) = w_{new}.
    About `zero_one_fraction` of the constraints are booleanity-style (b * b = b over a 0/1 input signal... realised
    as product of two 0/1 signals), which keeps ~90 % of the witness in {0,1} like a SHA/RSA bit-decomposed circuit.
    n_public signals (the first inputs) are public, as anon_aadhaar has 9.  Returns (zkey, wtns, vk)."""
    import random
    rnd = random.Random(rng_seed)
    n_vars = 1 + n_inputs + n_constraints
    power = (n_constraints + n_public + 1 - 1).bit_length()
    N = 1 << power
    tw = toxic(seed)
    tau, alpha, beta, gamma, delta = (tw[k] for k in ("tau", "alpha", "beta", "gamma", "delta"))
    dinv, ginv = pow(delta, -1, R), pow(gamma, -1, R)
    # witness + sparse rows
    w = [1] + [rnd.randrange(2) if rnd.random() < zero_one_fraction else rnd.randrange(R) for _ in range(n_inputs)]
    A_rows, B_rows, C_sig = [], [], []
    bits = [i for i in range(1, n_inputs + 1) if w[i] in (0, 1)] or [1]
    for k in range(n_constraints):
        new = 1 + n_inputs + k
        if rnd.random() < zero_one_fraction:
            i, j = rnd.choice(bits), rnd.choice(bits)
            ra, rb = [(i, 1)], [(j, 1)]
        else:
            hi = new
            ra = [(rnd.randrange(hi), rnd.randrange(1, 1 << 16)) for _ in range(rnd.randint(1, max_terms))]
            rb = [(rnd.randrange(hi), rnd.randrange(1, 1 << 16)) for _ in range(rnd.randint(1, max_terms))]
        va = sum(c * w[s] for s, c in ra) % R
        vb = sum(c * w[s] for s, c in rb) % R
        w.append(va * vb % R)
        if w[-1] in (0, 1):
            bits.append(new)
        A_rows.append(ra)
        B_rows.append(rb)
        C_sig.append(new)
    assert len(w) == n_vars

    backend.ntt_release_domain()
    backend.ntt_init_domain(backend.get_root_of_unity(2 * N))
    try:
        L = words_to_ints(lagrange_at_tau(backend, tau, power))
        L2 = lagrange_at_tau(backend, tau, power + 1)
    finally:
        backend.ntt_release_domain()
    u, v, wv = [0] * n_vars, [0] * n_vars, [0] * n_vars
    recs = []
    for k in range(n_constraints):
        for s, c in A_rows[k]:
            u[s] = (u[s] + c * L[k]) % R
            recs.append((0, k, s, c))
        for s, c in B_rows[k]:
            v[s] = (v[s] + c * L[k]) % R
            recs.append((1, k, s, c))
        wv[C_sig[k]] = (wv[C_sig[k]] + L[k]) % R
    for i in range(n_public + 1):  # snarkjs public-input rows
        u[i] = (u[i] + L[n_constraints + i]) % R
        recs.append((0, n_constraints + i, i, 1))
    comb = [(beta * u[s] + alpha * v[s] + wv[s]) % R for s in range(n_vars)]
    pA = fixed_base(backend, ints_to_words(u))
    pB1 = fixed_base(backend, ints_to_words(v))
    pB2 = fixed_base(backend, ints_to_words(v), g2=True)
    pC = fixed_base(backend, ints_to_words([x * dinv % R for x in comb[n_public + 1:]]))
    pIC = fixed_base(backend, ints_to_words([x * ginv % R for x in comb[:n_public + 1]]))
    pH = fixed_base(backend, backend.scalar_mul_vec(word(dinv), np.ascontiguousarray(L2[1::2])))
    vk1 = fixed_base(backend, ints_to_words([alpha, beta, delta]))
    vk2 = fixed_base(backend, ints_to_words([beta, gamma, delta]), g2=True)
    r2 = MONT_R * MONT_R % R
    sec4 = struct.pack("<I", len(recs)) + b"".join(
        struct.pack("<III", m, row, s) + (c * r2 % R).to_bytes(32, "little") for m, row, s, c in recs)
    header = (struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little") +
              struct.pack("<III", n_vars, n_public, N) + vk1[0].tobytes() + vk1[1].tobytes() + vk2[0].tobytes() +
              vk2[1].tobytes() + vk1[2].tobytes() + vk2[2].tobytes())
    sections = [(1, struct.pack("<I", 1)), (2, header), (3, pIC.tobytes()), (4, sec4), (5, pA.tobytes()), (6, pB1.tobytes()),
                (7, pB2.tobytes()), (8, pC.tobytes()), (9, pH.tobytes())]
    zkey = b"zkey" + struct.pack("<II", 1, len(sections)) + b"".join(_section(i, p) for i, p in sections)
    wt_hdr = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", n_vars)
    wtns = (b"wtns" + struct.pack("<II", 2, 2) + _section(1, wt_hdr) + _section(2, b"".join(x.to_bytes(32, "little") for x in w)))
    std1 = backend.convert_montgomery(vk1, False, kind="affine")
    std2 = backend.convert_montgomery(vk2, False, kind="g2_affine")
    vk = dict(alpha1=std1[0], beta2=std2[0], gamma2=std2[1], delta2=std2[2],
              ic=backend.convert_montgomery(pIC, False, kind="affine"), n_public=n_public)
    return zkey, wtns, vk


def vk_json(vk) -> str:
    """verification_key.json as snarkjs lays it out (the fields /root/reference/src/cache.rs:84-108 reads:
    vk_alpha_1, vk_beta_2, vk_gamma_2, vk_delta_2, IC, nPublic; snarkjs's vk_alphabeta_12 is not used there)."""
    import json

    def dec(w):
        return str(int.from_bytes(np.ascontiguousarray(w, dtype=np.uint32).tobytes(), "little"))

    def g1(p):
        p = np.asarray(p).reshape(-1)
        return [dec(p[:8]), dec(p[8:16]), "1"]

    def g2(p):
        p = np.asarray(p).reshape(-1)
        return [[dec(p[:8]), dec(p[8:16])], [dec(p[16:24]), dec(p[24:32])], ["1", "0"]]

    return json.dumps({"protocol": "groth16", "curve": "bn128", "nPublic": int(vk["n_public"]),
                       "vk_alpha_1": g1(vk["alpha1"]), "vk_beta_2": g2(vk["beta2"]), "vk_gamma_2": g2(vk["gamma2"]),
                       "vk_delta_2": g2(vk["delta2"]), "IC": [g1(p) for p in vk["ic"]]}, indent=1)


if __name__ == "__main__":
    # python tools/synth.py --constraints N [--out DIR]: generate ComplexCircuit(N, N) with the product library on cuda:0
    # and leave circuit.zkey / witness.wtns / vk.npz / verification_key.json in bench.py's instance cache (used by
    # `bench.py --impl reference`, whose own process never loads the product library).
    import argparse
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    ap = argparse.ArgumentParser()
    ap.add_argument("--constraints", type=int, required=True)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import __graft_entry__ as ge
    import bench
    if a.out:
        bench.INSTANCE_DIR = os.path.dirname(os.path.abspath(a.out))
    pkg = ge.load_package()
    lib = pkg.lib()
    lib.set_device("CUDA", 0)
    bench.save_instance(a.constraints, *make_complex_circuit(lib, a.constraints))
    print("instance written to", bench.instance_paths(a.constraints)[0])
