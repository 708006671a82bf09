"""Pins the pure-Python restatement (oracle/bn254_py.py) against the reference's own C++ compiled here
(oracle/_ref) and against the constants quoted in the reference's sources."""
import numpy as np

from oracle import bn254_py as O
from util import R, Q, array_to_ints, from_words, g1_points_multiples, g2_points_multiples, ints_to_array, rand_scalars, to_words

import icicle_snark_b200 as pkg
B = pkg.bindings


def test_constants_from_reference_sources(ref):
    # W[28] == rou, W[k]^2 == W[k-1] (src/cache.rs:25-54); omega(logn) via get_root_of_unity (ntt.cpp:54-63)
    assert pow(O.ROU_2_28, 1 << 28, R) == 1 and pow(O.ROU_2_28, 1 << 27, R) != 1
    for logn in (1, 3, 17, 22, 28):
        assert from_words(ref.get_root_of_unity(1 << logn)) == O.omega(logn)
    assert from_words(ref.get_root_of_unity(100002)) == O.omega(17)  # ceil(log2)
    g = ref.to_affine(ref.generator())
    assert [from_words(g[:8]), from_words(g[8:])] == list(O.G1_GEN)
    g2 = ref.to_affine(ref.generator(g2=True), g2=True)
    assert [from_words(g2[8 * i:8 * i + 8]) for i in range(4)] == [O.G2_GEN[0][0], O.G2_GEN[0][1], O.G2_GEN[1][0], O.G2_GEN[1][1]]
    assert O.G1.is_on_curve(O.G1_GEN) and O.G2.is_on_curve(O.G2.gen)


def test_field_ops_match_reference(ref, rng):
    a, av = rand_scalars(rng, 64)
    b, bv = rand_scalars(rng, 64)
    for i in range(64):
        assert from_words(ref.fr_mul(a[i], b[i])) == av[i] * bv[i] % R
        assert from_words(ref.fr_add(a[i], b[i])) == (av[i] + bv[i]) % R
        assert from_words(ref.fr_sub(a[i], b[i])) == (av[i] - bv[i]) % R
    assert from_words(ref.fr_inv(a[0])) == pow(av[0], -1, R)


def test_msm_g1_matches_reference(ref, rng):
    n = 300
    pts, aff = g1_points_multiples(n, start=5)
    pts[7] = 0
    aff[7] = None  # point at infinity (0,0)
    sc, sv = rand_scalars(rng, n)
    sv[3] = 0
    sc[3] = 0
    got = O.g1_projective_words_to_affine(list(ref.msm(sc, pts)[0]))
    want = O.G1.msm(sv, aff)
    assert got == want
    # known-dlog KAT: sum s_i (5+i) G
    k = sum(s * (5 + i) for i, s in enumerate(sv) if i != 7) % R
    assert want == O.G1.mul(O.G1_GEN, k)


def test_msm_g2_matches_reference(ref, rng):
    n = 40
    pts, aff = g2_points_multiples(n, start=3)
    sc, sv = rand_scalars(rng, n)
    got = O.g2_projective_words_to_affine(list(ref.msm(sc, pts, g2=True)[0]))
    assert got == O.G2.msm(sv, aff)


def test_ntt_matches_reference(ref, rng):
    ref.ntt_init_domain(ref.get_root_of_unity(1 << 10))
    try:
        for logn in (1, 4, 7):
            x, xv = rand_scalars(rng, 1 << logn)
            assert array_to_ints(ref.ntt(x, B.kForward)) == O.ntt(xv)
            assert array_to_ints(ref.ntt(x, B.kInverse)) == O.ntt(xv, inverse=True)
        xv = [rng.integers(0, 1 << 62).item() for _ in range(16)]
        assert O.ntt(xv) == O.ntt_naive(xv)
        # coset forward = pre-multiply by g^j
        cfg = B.NTTConfig.default()
        g = 0x1234567
        for i, l in enumerate(to_words(g)):
            cfg.coset_gen[i] = int(l)
        x, xv = rand_scalars(rng, 32)
        assert array_to_ints(ref.ntt(x, B.kForward, cfg)) == O.ntt(xv, coset_gen=g)
        assert array_to_ints(ref.ntt(x, B.kInverse, cfg)) == O.ntt(xv, inverse=True, coset_gen=g)
    finally:
        ref.ntt_release_domain()


def test_montgomery_conversion_reference(ref, rng):
    x, xv = rand_scalars(rng, 16)
    m = ref.convert_montgomery(x, True)
    assert array_to_ints(m) == [v * O.MONT_R % R for v in xv]
    assert array_to_ints(ref.convert_montgomery(m, False)) == xv
