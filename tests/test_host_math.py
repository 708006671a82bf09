"""Host-side helpers of the product library (bn254_add/mul/inv, ecadd/ecsub/mul_scalar/to_affine, G2 twins;
the prover epilogue's only arithmetic) against the reference's compiled frontend. These run the SAME
field/curve templates (csrc/field.cuh, csrc/curve.cuh) the device kernels instantiate, on the host."""
import numpy as np
import pytest

from oracle import bn254_py as O
from util import R, Q, from_words, g1_points_multiples, g2_points_multiples, rand_scalars, to_words


def test_fr_helpers(lib, ref, rng):
    a, av = rand_scalars(rng, 200)
    b, bv = rand_scalars(rng, 200)
    edge = [0, 1, R - 1, R - 2, 2, (1 << 253)]
    for x in edge:
        for y in edge:
            xa, ya = to_words(x), to_words(y)
            assert from_words(lib.fr_mul(xa, ya)) == x * y % R
            assert from_words(lib.fr_add(xa, ya)) == (x + y) % R
            assert from_words(lib.fr_sub(xa, ya)) == (x - y) % R
    for i in range(200):
        assert np.array_equal(lib.fr_mul(a[i], b[i]), ref.fr_mul(a[i], b[i]))
        assert np.array_equal(lib.fr_add(a[i], b[i]), ref.fr_add(a[i], b[i]))
        assert np.array_equal(lib.fr_sub(a[i], b[i]), ref.fr_sub(a[i], b[i]))
    for i in range(8):
        assert np.array_equal(lib.fr_inv(a[i]), ref.fr_inv(a[i]))
        assert np.array_equal(lib.fr_pow(a[i], 12345 + i), ref.fr_pow(a[i], 12345 + i))
    assert from_words(lib.fr_inv(to_words(0))) == 0  # inverse(0) == 0 (modular_arithmetic.h:603)


def _curve_checks(lib, ref, rng, g2):
    gen = lib.generator(g2=g2)
    assert np.array_equal(lib.to_affine(gen, g2=g2), ref.to_affine(ref.generator(g2=g2), g2=g2))
    assert lib.is_on_curve(gen, g2=g2)
    sc, sv = rand_scalars(rng, 6)
    P = [lib.mul_scalar(gen, sc[i], g2=g2) for i in range(6)]
    Pr = [ref.mul_scalar(ref.generator(g2=g2), sc[i], g2=g2) for i in range(6)]
    for p, pr in zip(P, Pr):
        assert lib.is_on_curve(p, g2=g2) and ref.is_on_curve(p, g2=g2)
        assert ref.eq(p, pr, g2=g2) and lib.eq(p, pr, g2=g2)
        assert np.array_equal(lib.to_affine(p, g2=g2), ref.to_affine(pr, g2=g2))
    # add / sub / doubling / inverse pairs, vs the reference's complete formulas
    for i in range(5):
        s = lib.ecadd(P[i], P[i + 1], g2=g2)
        assert np.array_equal(lib.to_affine(s, g2=g2), ref.to_affine(ref.ecadd(Pr[i], Pr[i + 1], g2=g2), g2=g2))
        d = lib.ecsub(P[i], P[i + 1], g2=g2)
        assert np.array_equal(lib.to_affine(d, g2=g2), ref.to_affine(ref.ecsub(Pr[i], Pr[i + 1], g2=g2), g2=g2))
    dbl = lib.ecadd(P[0], P[0], g2=g2)  # P == Q branch
    assert np.array_equal(lib.to_affine(dbl, g2=g2), ref.to_affine(ref.ecadd(Pr[0], Pr[0], g2=g2), g2=g2))
    zero = lib.ecsub(P[0], P[0], g2=g2)  # P == -Q branch -> identity (0,y!=0,0)
    w = 16 if g2 else 8
    assert not zero[2 * w:].any() and not zero[:w].any() and zero[w:2 * w].any()
    assert not lib.to_affine(zero, g2=g2).any()  # to_affine(inf) == (0,0)
    assert np.array_equal(lib.to_affine(lib.ecadd(zero, P[1], g2=g2), g2=g2), lib.to_affine(P[1], g2=g2))
    assert lib.eq(zero, ref.ecsub(Pr[2], Pr[2], g2=g2), g2=g2)
    allzero = np.zeros_like(zero)
    assert not lib.eq(allzero, allzero, g2=g2) and not ref.eq(allzero, allzero, g2=g2)  # ffi_extern.cpp:9-16
    # from_affine round trip, incl. infinity
    a = lib.to_affine(P[3], g2=g2)
    assert lib.eq(lib.from_affine(a, g2=g2), P[3], g2=g2)
    assert lib.eq(lib.from_affine(np.zeros_like(a), g2=g2), zero, g2=g2)


def test_g1_helpers(lib, ref, rng):
    _curve_checks(lib, ref, rng, g2=False)


def test_g2_helpers(lib, ref, rng):
    _curve_checks(lib, ref, rng, g2=True)


def test_generate_points_are_valid(lib, ref):
    pts = lib.generate_affine_points(9)
    for p in pts:
        assert ref.is_on_curve(ref.from_affine(p))
    pts2 = lib.generate_affine_points(5, g2=True)
    for p in pts2:
        assert ref.is_on_curve(ref.from_affine(p, g2=True), g2=True)
    s = lib.generate_scalars(50)
    assert all(from_words(x) < R for x in s)


def test_python_oracle_agrees_with_host_helpers(lib, rng):
    sc, sv = rand_scalars(rng, 3)
    for i in range(3):
        got = O.g1_projective_words_to_affine(list(lib.mul_scalar(lib.generator(), sc[i])))
        assert got == O.G1.mul(O.G1_GEN, sv[i])
        got2 = O.g2_projective_words_to_affine(list(lib.mul_scalar(lib.generator(g2=True), sc[i], g2=True)))
        assert got2 == O.G2.mul(O.G2.gen, sv[i])


@pytest.mark.parametrize("fq", [0, 1])
def test_division_step_inverse_matches_fermat_on_host(lib, fq):
    """csrc/field_inv.cuh (the shared inversion of the batched affine accumulation): x * inverse_safegcd(x) == 1 for
    random Montgomery residues and the edge values 0 -> 0, 1, -1, raw 1, a power of two; equal to Fermat's inverse."""
    import ctypes as C
    import icicle_snark_b200 as pkg
    f = pkg.tools_lib().b200_inv_check
    f.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int]
    assert f(3000, 20261017 + fq, fq, 0) == 0


def test_epilogue_double_scalar_multiplication_on_host(lib):
    """csrc/host_math.h host_double_scalar_mul (s*pi_a + r*pi_b1 of the blinding epilogue, proof_helper.rs:274-295, with one
    shared doubling chain) == two plain double-and-add multiplications; zero / one scalars, equal, opposite and infinite points."""
    import ctypes as C
    import icicle_snark_b200 as pkg
    f = pkg.tools_lib().b200_double_mul_check
    f.argtypes = [C.c_int, C.c_uint64]
    assert f(40, 20261017) == 0
