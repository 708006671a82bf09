"""CPU models of two kernel-level restructurings, checked against the independent big-int oracle (oracle/bn254_py.py):

* csrc/msm_reduce_quad.cuh - the quad-cooperative group law: the product each lane computes in each step and what is
  exchanged, replayed with Python integers (four "lanes" that only see their own product and the shuffled results), for
  G1 and G2, including the exceptional cases; the result must be the oracle's affine sum / double.
* csrc/ntt.cu ntt_pass4_kernel - one radix-4 stage step (levels st and st-1 in registers) must equal two radix-2 DIF
  levels of the Stockham pass it replaces, index for index and twiddle for twiddle, for every k and every st.

These are models of the schedules (which operand pairs, which twiddle exponents), not of the arithmetic itself - the
GPU parity tests (tests/test_gpu_ops.py) run the real kernels against the reference.
"""
import random

import pytest

from oracle import bn254_py as O


# ------------------------------------------------------------------------------------------------ quad group law
class _Field:
    def __init__(self, ops, zero, one):
        self.ops, self.zero, self.one = ops, zero, one

    def mul(self, a, b): return self.ops.mul(a, b)
    def add(self, a, b): return self.ops.add(a, b)
    def sub(self, a, b): return self.ops.sub(a, b)
    def dbl(self, a): return self.ops.add(a, a)
    def is_zero(self, a): return self.ops.is_zero(a)
    def inv(self, a): return self.ops.inv(a)


FQ = _Field(O._FqOps, 0, 1)
FQ2 = _Field(O._Fq2Ops, O.Fq2(0, 0), O.Fq2(1, 0))


def _quad_step(F, xs, ys):
    """lane k multiplies xs[k] * ys[k]; every lane then sees all four results (width-4 shuffles)"""
    return [F.mul(xs[k], ys[k]) for k in range(4)]


def xyzz_add_quad(F, a, b):
    """mirror of xyzz_add_quad: a, b = (x, y, zz, zzz) or None for infinity"""
    if b is None:
        return a
    if a is None:
        return b
    ax, ay, azz, azzz = a
    bx, by, bzz, bzzz = b
    u1, u2, s1, s2 = _quad_step(F, [ax, bx, ay, by], [bzz, azz, bzzz, azzz])
    p, r = F.sub(u2, u1), F.sub(s2, s1)
    if F.is_zero(p):
        return xyzz_dbl_quad(F, a) if F.is_zero(r) else None
    pp, rr, zzA, zzzA = _quad_step(F, [p, r, azz, azzz], [p, r, bzz, bzzz])
    ppp, q, zz3, _ = _quad_step(F, [p, u1, zzA, zzA], [pp, pp, pp, pp])
    x3 = F.sub(F.sub(rr, ppp), F.dbl(q))
    t1, t2, zzz3, _ = _quad_step(F, [r, s1, zzzA, zzzA], [F.sub(q, x3), ppp, ppp, ppp])
    return (x3, F.sub(t1, t2), zz3, zzz3)


def xyzz_dbl_quad(F, a):
    if a is None:
        return None
    ax, ay, azz, azzz = a
    if F.is_zero(ay):
        return None
    u = F.dbl(ay)
    v, x2, _, _ = _quad_step(F, [u, ax, u, ax], [u, ax, u, ax])
    m = F.add(F.dbl(x2), x2)
    w, s, mm, zz3 = _quad_step(F, [u, ax, m, v], [v, v, m, azz])
    x3 = F.sub(mm, F.dbl(s))
    t1, t2, zzz3, _ = _quad_step(F, [m, w, w, w], [F.sub(s, x3), ay, azzz, azzz])
    return (x3, F.sub(t1, t2), zz3, zzz3)


def _to_affine(F, p):
    if p is None or F.is_zero(p[2]):
        return None
    x, y, zz, zzz = p
    return (F.mul(x, F.inv(zz)), F.mul(y, F.inv(zzz)))


def _from_affine(F, P, rng, scale=True):
    """affine -> XYZZ with a random Z (x Z^2, y Z^3, Z^2, Z^3) so that the general-coordinates path is exercised"""
    if P is None:
        return None
    z = F.one
    if scale:
        z = rng.randrange(2, O.Q_MOD) if F is FQ else O.Fq2(rng.randrange(2, O.Q_MOD), rng.randrange(1, O.Q_MOD))
    z2 = F.mul(z, z)
    z3 = F.mul(z2, z)
    return (F.mul(P[0], z2), F.mul(P[1], z3), z2, z3)


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_quad_group_law_schedule_matches_the_oracle(group):
    rng = random.Random(20261017)
    F, curve, gen = (FQ, O.G1, O.G1_GEN) if group == "g1" else (FQ2, O.G2, O.G2.gen)
    pts = [curve.mul(gen, rng.randrange(1, O.R_MOD)) for _ in range(4)]
    P, Q = pts[0], pts[1]
    cases = [(P, Q), (Q, P), (P, P), (P, curve.neg(P)), (None, Q), (P, None), (None, None), (pts[2], pts[3])]
    for A, B in cases:
        for scale in (True, False):
            got = _to_affine(F, xyzz_add_quad(F, _from_affine(F, A, rng, scale), _from_affine(F, B, rng, scale)))
            assert got == curve.add(A, B), (group, A is None, B is None, scale)
    for A in (P, Q, None):
        got = _to_affine(F, xyzz_dbl_quad(F, _from_affine(F, A, rng)))
        assert got == curve.add(A, A)
    # a running sum and a double-and-add built from the two primitives (what the level-1 kernel does)
    run = None
    want = None
    for X in pts:
        run = xyzz_add_quad(F, run, _from_affine(F, X, rng))
        want = curve.add(want, X)
    assert _to_affine(F, run) == want
    k = 0b1011010011
    acc = None
    for bit in bin(k)[2:]:
        acc = xyzz_dbl_quad(F, acc)
        if bit == "1":
            acc = xyzz_add_quad(F, acc, run)
    assert _to_affine(F, acc) == curve.mul(want, k)


# ------------------------------------------------------------------------------------------------ radix-4 stage steps
def _radix2_levels(vals, k, itw, p):
    """the stage loop of ntt_pass_kernel (radix-2 DIF levels st = k-1 .. 0) on one column of R = 2^k values"""
    R = 1 << k
    s = list(vals)
    for st in range(k - 1, -1, -1):
        half = 1 << st
        for bb in range(R // 2):
            lo = bb & (half - 1)
            t = ((bb >> st) << (st + 1)) | lo
            x, y = s[t], s[t + half]
            d = (x - y) % p
            if st > 0 and lo:
                d = d * itw[lo << (k - 1 - st)] % p
            s[t], s[t + half] = (x + y) % p, d
    return s


def _radix4_steps(vals, k, itw, p):
    """the stage loop of ntt_pass4_kernel: an odd k starts with one radix-2 level, then two levels per step"""
    R = 1 << k
    s = list(vals)
    st = k - 1
    if k & 1:
        half = 1 << st
        for bb in range(R // 2):
            lo = bb & (half - 1)
            t = ((bb >> st) << (st + 1)) | lo
            x, y = s[t], s[t + half]
            d = (x - y) % p
            if st > 0 and lo:
                d = d * itw[lo << (k - 1 - st)] % p
            s[t], s[t + half] = (x + y) % p, d
        st -= 1
    while st >= 1:
        half, quarter = 1 << st, 1 << (st - 1)
        for bb in range(R // 4):
            lo = bb & (quarter - 1)
            t0 = ((bb >> (st - 1)) << (st + 1)) | lo
            i0, i1, i2, i3 = t0, t0 + quarter, t0 + half, t0 + half + quarter
            x0, x1, x2, x3 = s[i0], s[i1], s[i2], s[i3]
            e0 = lo << (k - 1 - st)
            a0, a2, a1, a3 = (x0 + x2) % p, (x0 - x2) % p, (x1 + x3) % p, (x1 - x3) % p
            if e0:
                a2 = a2 * itw[e0] % p
            a3 = a3 * itw[e0 + (R >> 2)] % p
            b0, b1, b2, b3 = (a0 + a1) % p, (a0 - a1) % p, (a2 + a3) % p, (a2 - a3) % p
            if e0:
                w2 = itw[2 * e0]
                b1, b3 = b1 * w2 % p, b3 * w2 % p
            s[i0], s[i1], s[i2], s[i3] = b0, b1, b2, b3
        st -= 2
    return s


@pytest.mark.parametrize("k", list(range(0, 9)))
def test_radix4_stage_steps_equal_two_radix2_levels(k):
    p = O.R_MOD
    rng = random.Random(100 + k)
    R = 1 << k
    w = O.omega(k) if k > 0 else 1
    itw = [pow(w, i, p) for i in range(max(R // 2, 1))]
    vals = [rng.randrange(p) for _ in range(R)]
    want = _radix2_levels(vals, k, itw, p)
    assert _radix4_steps(vals, k, itw, p) == want
    if k:
        # and both are the DFT of the column in bit-reversed order (what the pass kernel un-reverses on its way out)
        dft = O.ntt_naive(vals)
        rev = [int(format(i, f"0{k}b")[::-1], 2) for i in range(R)]
        assert [want[rev[i]] for i in range(R)] == dft
