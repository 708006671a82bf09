"""CPU: the library's host pairing and Groth16 verification (csrc/pairing.cu) against the reference's own
`bn254_pairing` / `bn254_pairing_target_field_*` compiled into oracle/_ref, the reference's bilinearity test
(wrappers/rust/icicle-core/src/pairing/tests.rs:8-28), and the verification equation (src/proof_helper.rs:319-372)
on the committed golden proofs.  Pairing is host code in the reference as well; nothing here needs a GPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import icicle_snark_b200 as pkg
from oracle import groth16_ref as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def load_vk(n):
    vkz = np.load(os.path.join(GOLD, f"complex_{n}.vk.npz"))
    vk = {k: vkz[k] for k in ("alpha1", "beta2", "gamma2", "delta2", "ic")}
    vk["n_public"] = int(vkz["n_public"])
    return vk


def proof_points(path):
    d = json.load(open(path))
    w = lambda s: np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint32)
    return {"pi_a": np.concatenate([w(d["pi_a"][0]), w(d["pi_a"][1])]),
            "pi_b": np.concatenate([w(d["pi_b"][0][0]), w(d["pi_b"][0][1]), w(d["pi_b"][1][0]), w(d["pi_b"][1][1])]),
            "pi_c": np.concatenate([w(d["pi_c"][0]), w(d["pi_c"][1])])}


def test_pairing_is_bit_exact_with_the_reference(lib, ref):
    ps, qs = ref.generate_affine_points(6), ref.generate_affine_points(6, g2=True)
    for p, q in zip(ps, qs):
        assert np.array_equal(lib.pairing(p, q), ref.pairing(p, q))
    g1, g2 = lib.to_affine(lib.generator()), lib.to_affine(lib.generator(g2=True), g2=True)
    e = lib.pairing(g1, g2)
    assert np.array_equal(e, ref.pairing(g1, g2)) and not np.array_equal(e, lib.target_from_u32(1))
    # the reference does not special-case the affine zero point; neither side may diverge on it
    z1, z2 = np.zeros(16, dtype=np.uint32), np.zeros(32, dtype=np.uint32)
    for p, q in ((z1, qs[0]), (ps[0], z2), (z1, z2)):
        assert np.array_equal(lib.pairing(p, q), ref.pairing(p, q))


def test_bilinearity_as_the_reference_tests_it(lib):
    p, q = lib.generate_affine_points(1)[0], lib.generate_affine_points(1, g2=True)[0]
    s = np.zeros(8, dtype=np.uint32)
    s[0] = 42
    ps = lib.to_affine(lib.mul_scalar(lib.from_affine(p), s))
    qs = lib.to_affine(lib.mul_scalar(lib.from_affine(q, g2=True), s, g2=True), g2=True)
    e = lib.pairing(p, q)
    assert np.array_equal(lib.pairing(ps, q), lib.pairing(p, qs))
    assert np.array_equal(lib.pairing(ps, q), lib.target_pow(e, 42))
    # e(P,Q) * e(-P,Q) == 1
    neg = lib.to_affine(lib.ecsub(lib.ecsub(lib.from_affine(p), lib.from_affine(p)), lib.from_affine(p)))
    assert np.array_equal(lib.target_mul(e, lib.pairing(neg, q)), lib.target_from_u32(1))


def test_target_field_helpers_match_the_reference(lib, ref):
    xs = lib.target_generate(4)
    for x in xs:  # canonical residues
        assert all(int.from_bytes(x[8 * k:8 * k + 8].tobytes(), "little") < P_MOD for k in range(12))
    a, b = xs[0], xs[1]
    assert np.array_equal(lib.target_add(a, b), ref.target_add(a, b))
    assert np.array_equal(lib.target_sub(a, b), ref.target_sub(a, b))
    assert np.array_equal(lib.target_mul(a, b), ref.target_mul(a, b))
    assert np.array_equal(lib.target_inv(a), ref.target_inv(a))
    assert np.array_equal(lib.target_mul(a, lib.target_inv(a)), lib.target_from_u32(1))
    for e in (0, 1, 2, 77, 65537):
        assert np.array_equal(lib.target_pow(b, e), ref.target_pow(b, e))
    assert np.array_equal(lib.target_from_u32(12345), ref.target_from_u32(12345))


@pytest.mark.parametrize("n", [6, 100])
def test_verify_equation_on_golden_proofs(lib, ref, n):
    vk = load_vk(n)
    base = os.path.join(GOLD, f"complex_{n}")
    public = [int(x) for x in json.load(open(base + ".public.json"))]
    p11, prs = proof_points(base + ".proof_r1s1.json"), proof_points(base + ".proof_rs.json")
    for proof in (p11, prs):
        assert pkg.groth16_verify_points(lib, proof, public, vk) and G.verify(ref, proof, public, vk)
    bad_public = [public[0] + 1] + public[1:]
    mixed = dict(p11, pi_c=prs["pi_c"])
    swapped = dict(p11, pi_a=p11["pi_c"])
    for proof, pub in ((p11, bad_public), (mixed, public), (swapped, public)):
        assert not pkg.groth16_verify_points(lib, proof, pub, vk) and not G.verify(ref, proof, pub, vk)
    with pytest.raises(ValueError):
        pkg.groth16_verify_points(lib, p11, public + [1], vk)


def test_verify_files_mirror(lib, tmp_path):
    base = os.path.join(GOLD, "complex_100")
    pkg.groth16_verify(base + ".proof_r1s1.json", base + ".public.json", base + ".vk.json", lib=lib)
    pkg.groth16_verify(base + ".proof_rs.json", base + ".public.json", base + ".vk.json", lib=lib)
    # a proof for another circuit / a tampered public input: the Rust code's assert!(..., "Verification failed")
    other = os.path.join(GOLD, "complex_6")
    with pytest.raises(AssertionError, match="Verification failed"):
        pkg.groth16_verify(other + ".proof_r1s1.json", base + ".public.json", base + ".vk.json", lib=lib)
    pub = json.load(open(base + ".public.json"))
    pub[0] = str(int(pub[0]) + 1)
    bad = tmp_path / "public.json"
    bad.write_text(json.dumps(pub))
    with pytest.raises(AssertionError, match="Verification failed"):
        pkg.groth16_verify(base + ".proof_r1s1.json", str(bad), base + ".vk.json", lib=lib)
    # unreadable or malformed inputs are errors, not "invalid proof"
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(str(tmp_path / "missing.json"), base + ".public.json", base + ".vk.json", lib=lib)
    trunc = tmp_path / "vk.json"
    trunc.write_text(open(base + ".vk.json").read()[:200])
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(base + ".proof_r1s1.json", base + ".public.json", str(trunc), lib=lib)
    novk = tmp_path / "vk2.json"
    d = json.load(open(base + ".vk.json"))
    del d["vk_gamma_2"]
    novk.write_text(json.dumps(d))
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(base + ".proof_r1s1.json", base + ".public.json", str(novk), lib=lib)
    # serde's typing: coordinates are strings (a bare JSON number is an error), nothing may follow the document
    nums = tmp_path / "public_numbers.json"
    nums.write_text("[" + ", ".join(json.load(open(base + ".public.json"))) + "]")
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(base + ".proof_r1s1.json", str(nums), base + ".vk.json", lib=lib)
    trailing = tmp_path / "proof_trailing.json"
    trailing.write_text(open(base + ".proof_r1s1.json").read() + " x")
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(str(trailing), base + ".public.json", base + ".vk.json", lib=lib)
    # snarkjs's vk_alphabeta_12 is ignored; the number of public values must equal nPublic and IC must hold nPublic + 1
    # points (the reference's `public.iter().take(n_public)` zip would read missing inputs as absent terms)
    d = json.load(open(base + ".vk.json"))
    d["vk_alphabeta_12"] = [[["1", "2"], ["3", "4"], ["5", "6"]], [["7", "8"], ["9", "10"], ["11", "12"]]]
    vk2 = tmp_path / "vk_snarkjs.json"
    vk2.write_text(json.dumps(d, indent=1))
    pkg.groth16_verify(base + ".proof_r1s1.json", base + ".public.json", str(vk2), lib=lib)
    extra = tmp_path / "public_extra.json"
    extra.write_text(json.dumps(json.load(open(base + ".public.json")) + ["5"]))
    short = tmp_path / "public_short.json"
    short.write_text("[]")
    for bad_pub in (extra, short):
        with pytest.raises(pkg.IcicleError):
            pkg.groth16_verify(base + ".proof_r1s1.json", str(bad_pub), str(vk2), lib=lib)
    d2 = dict(d, IC=d["IC"] + [d["IC"][0]])
    vk3 = tmp_path / "vk_long_ic.json"
    vk3.write_text(json.dumps(d2))
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_verify(base + ".proof_r1s1.json", base + ".public.json", str(vk3), lib=lib)
    ok = C.c_int(7)
    assert lib.dll.b200_groth16_verify_files(None, None, None, C.byref(ok)) == pkg.ERRORS.index("INVALID_POINTER")


R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_verifier_rejects_non_canonical_and_off_curve_inputs(lib):
    """Aliased public inputs (x + r), non-canonical coordinates (x + q), points off the curve and G2 points outside
    the order-r subgroup must not verify (snarkjs rejects them; the reference's verifier does not look)."""
    n = 6
    vk = load_vk(n)
    base = os.path.join(GOLD, f"complex_{n}")
    public = [int(x) for x in json.load(open(base + ".public.json"))]
    good = proof_points(base + ".proof_rs.json")
    assert pkg.groth16_verify_points(lib, good, public, vk)
    words = lambda v: np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint32).copy()
    val = lambda w: int.from_bytes(np.ascontiguousarray(w, dtype=np.uint32).tobytes(), "little")
    # public input aliasing: public[0] + r is the same field element
    assert public[0] + R_MOD < 1 << 256
    assert not pkg.groth16_verify_points(lib, good, [public[0] + R_MOD] + public[1:], vk)
    # coordinate malleability: pi_a.x + q
    bad = {k: v.copy() for k, v in good.items()}
    bad["pi_a"][:8] = words(val(good["pi_a"][:8]) + P_MOD)
    assert not pkg.groth16_verify_points(lib, bad, public, vk)
    bad = {k: v.copy() for k, v in good.items()}
    bad["pi_b"][8:16] = words(val(good["pi_b"][8:16]) + P_MOD)
    assert not pkg.groth16_verify_points(lib, bad, public, vk)
    # off the curve: y + 1
    for name, off in (("pi_a", 8), ("pi_c", 8), ("pi_b", 16)):
        bad = {k: v.copy() for k, v in good.items()}
        bad[name][off:off + 8] = words((val(good[name][off:off + 8]) + 1) % P_MOD)
        assert not pkg.groth16_verify_points(lib, bad, public, vk)
    # on the twist but outside the order-r subgroup (the twist's cofactor is > 1, so almost every twist point is)
    from oracle import bn254_py as O
    Q = O.Q_MOD

    def fq2_sqrt(a):  # q = 3 mod 4
        if a.is_zero():
            return a
        norm = (a.c0 * a.c0 + a.c1 * a.c1) % Q
        alpha = pow(norm, (Q + 1) // 4, Q)
        if alpha * alpha % Q != norm:
            return None
        for sgn in (1, -1):
            delta = (a.c0 + sgn * alpha) * pow(2, -1, Q) % Q
            x0 = pow(delta, (Q + 1) // 4, Q)
            if x0 * x0 % Q == delta and x0:
                r = O.Fq2(x0, a.c1 * pow(2 * x0, -1, Q))
                if r * r == a:
                    return r
        return None

    def mul_no_reduce(P, k):
        acc = None
        while k:
            if k & 1:
                acc = O.G2.add(acc, P)
            P = O.G2.add(P, P)
            k >>= 1
        return acc

    found = None
    for x0 in range(1, 50):
        x = O.Fq2(x0, 1)
        y = fq2_sqrt(x * x * x + O.G2.b)
        if y is not None and O.G2.is_on_curve((x, y)) and mul_no_reduce((x, y), R_MOD) is not None:
            found = (x, y)
            break
    assert found is not None
    assert mul_no_reduce(O.G2.gen, R_MOD) is None  # sanity of the helper: the generator has order r
    bad = {k: v.copy() for k, v in good.items()}
    bad["pi_b"] = np.concatenate([words(found[0].c0), words(found[0].c1), words(found[1].c0), words(found[1].c1)])
    assert not pkg.groth16_verify_points(lib, bad, public, vk)
    # ... while a genuine subgroup point in pi_b's place passes the input checks and only fails the pairing equation
    g2w = np.array(O.g2_affine_to_words(O.G2.gen), dtype=np.uint32)
    bad["pi_b"] = g2w
    assert not pkg.groth16_verify_points(lib, bad, public, vk)
