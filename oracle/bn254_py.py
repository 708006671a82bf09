"""Pure-Python big-int model of the BN254 arithmetic on the Groth16 proving path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this file; the product (icicle-snark_b200/) never does.

This is an *independent* restatement (Python ints, affine coordinates, textbook
formulas) of what the reference computes with limb arithmetic:

  * Fr / Fq constants      /root/reference/icicle/include/icicle/fields/snark_fields/bn254_scalar.h:9-10,68-69
                           /root/reference/icicle/include/icicle/fields/snark_fields/bn254_base.h:8-9
  * generators, b, b'      /root/reference/icicle/include/icicle/curves/params/bn254.h:19-52
  * 2-adic root table W    /root/reference/src/cache.rs:25-54   (W[28] == rou, W[k]^2 == W[k-1])
  * omega(logn)            /root/reference/icicle/include/icicle/math/modular_arithmetic.h:61-73
  * NTT conventions        /root/reference/icicle/backend/cpu/include/ntt_cpu.h (kNN, inverse scales by 1/N)
  * MSM                    /root/reference/icicle/backend/cpu/src/curve/cpu_msm.hpp:41-454 (group result only)
  * Montgomery R = 2^256   /root/reference/icicle/include/icicle/fields/params_gen.h:35-50

Pinned (tests/test_oracle.py) against oracle/_ref (the reference's own C++ compiled
here) and against the constants quoted above.  Loops are Python loops: small cases only.
"""
from __future__ import annotations

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
MONT_R = 1 << 256
ROU_2_28 = 0x2A3C09F0A58A7E8500E0A7EB8EF62ABC402D111E41112ED49BD61B6E725B19F0
TWO_ADICITY = 28
G1_GEN = (1, 2)
G2_GEN = (
    (0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED,
     0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
    (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA,
     0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B),
)
G1_B = 3
G2_B = (0x2B149D40CEB8AAAE81BE18991BE06AC3B5B4C5E559DBEFA33267E6DC24A138E5,
        0x009713B03AF0FED4CD2CAFADEED8FDF4A74FA084E52D1852E4A2BD0685C315D2)


# ----------------------------------------------------------------------------- limbs
def to_limbs(x: int, n: int = 8) -> list[int]:
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_limbs(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (32 * i)
    return v


def to_bytes32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def from_bytes32(b: bytes) -> int:
    return int.from_bytes(b, "little")


# ----------------------------------------------------------------------------- Fr helpers
def omega(logn: int) -> int:
    """Primitive 2^logn-th root of unity, as modular_arithmetic.h:61-73 derives it."""
    assert 0 <= logn <= TWO_ADICITY
    w = ROU_2_28
    for _ in range(TWO_ADICITY - logn):
        w = w * w % R_MOD
    return w


def finv(x: int, p: int) -> int:
    """inverse(0) == 0, as modular_arithmetic.h:603."""
    return pow(x, p - 2, p) if x % p else 0


def ntt(values: list[int], inverse: bool = False, coset_gen: int = 1) -> list[int]:
    """Natural-order in, natural-order out. forward: out[k] = sum_j in[j] g^j w^(jk);
    inverse: out[j] = g^-j / N * sum_k in[k] w^(-jk)."""
    n = len(values)
    logn = n.bit_length() - 1
    assert 1 << logn == n
    w = omega(logn)
    if inverse:
        w = finv(w, R_MOD)
    a = [v % R_MOD for v in values]
    if not inverse and coset_gen != 1:
        g = 1
        for j in range(n):
            a[j] = a[j] * g % R_MOD
            g = g * coset_gen % R_MOD
    # iterative radix-2 (bit-reverse then DIT)
    rev = [0] * n
    for i in range(n):
        rev[i] = (rev[i >> 1] >> 1) | ((i & 1) << (logn - 1)) if logn else 0
    a = [a[rev[i]] for i in range(n)]
    length = 2
    while length <= n:
        wl = pow(w, n // length, R_MOD)
        for s in range(0, n, length):
            t = 1
            for j in range(length // 2):
                u, v = a[s + j], a[s + j + length // 2] * t % R_MOD
                a[s + j] = (u + v) % R_MOD
                a[s + j + length // 2] = (u - v) % R_MOD
                t = t * wl % R_MOD
        length <<= 1
    if inverse:
        ninv = finv(n, R_MOD)
        a = [x * ninv % R_MOD for x in a]
        if coset_gen != 1:
            gi = finv(coset_gen, R_MOD)
            g = 1
            for j in range(n):
                a[j] = a[j] * g % R_MOD
                g = g * gi % R_MOD
    return a


def ntt_naive(values: list[int], inverse: bool = False) -> list[int]:
    """O(n^2) definition, used to pin ntt() itself."""
    n = len(values)
    w = omega(n.bit_length() - 1)
    if inverse:
        w = finv(w, R_MOD)
    out = [sum(values[j] * pow(w, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]
    if inverse:
        ninv = finv(n, R_MOD)
        out = [x * ninv % R_MOD for x in out]
    return out


# ----------------------------------------------------------------------------- Fq2 (u^2 = -1)
class Fq2:
    __slots__ = ("c0", "c1")

    def __init__(self, c0=0, c1=0):
        self.c0 = c0 % Q_MOD
        self.c1 = c1 % Q_MOD

    def __add__(self, o):
        return Fq2(self.c0 + o.c0, self.c1 + o.c1)

    def __sub__(self, o):
        return Fq2(self.c0 - o.c0, self.c1 - o.c1)

    def __neg__(self):
        return Fq2(-self.c0, -self.c1)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq2(self.c0 * o, self.c1 * o)
        return Fq2(self.c0 * o.c0 - self.c1 * o.c1, self.c0 * o.c1 + self.c1 * o.c0)

    def __eq__(self, o):
        return self.c0 == o.c0 and self.c1 == o.c1

    def is_zero(self):
        return self.c0 == 0 and self.c1 == 0

    def inv(self):
        d = finv(self.c0 * self.c0 + self.c1 * self.c1, Q_MOD)
        return Fq2(self.c0 * d, -self.c1 * d)

    def __repr__(self):
        return f"Fq2({hex(self.c0)}, {hex(self.c1)})"


class _FqOps:
    zero = 0
    one = 1

    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def mul(a, b): return a * b % Q_MOD
    @staticmethod
    def neg(a): return (-a) % Q_MOD
    @staticmethod
    def inv(a): return finv(a, Q_MOD)
    @staticmethod
    def is_zero(a): return a % Q_MOD == 0


class _Fq2Ops:
    zero = Fq2(0, 0)
    one = Fq2(1, 0)

    @staticmethod
    def add(a, b): return a + b
    @staticmethod
    def sub(a, b): return a - b
    @staticmethod
    def mul(a, b): return a * b
    @staticmethod
    def neg(a): return -a
    @staticmethod
    def inv(a): return a.inv()
    @staticmethod
    def is_zero(a): return a.is_zero()


# ----------------------------------------------------------------------------- affine EC (None == infinity)
class Curve:
    def __init__(self, F, b, gen):
        self.F, self.b, self.gen = F, b, gen

    def is_on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.is_zero(F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if F.is_zero(F.sub(x1, x2)):
            if F.is_zero(F.add(y1, y2)):
                return None
            lam = F.mul(F.mul(F.mul(x1, x1), F.add(F.add(F.one, F.one), F.one)), F.inv(F.add(y1, y1)))
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def mul(self, P, k: int):
        k %= R_MOD
        acc = None
        while k:
            if k & 1:
                acc = self.add(acc, P)
            P = self.add(P, P)
            k >>= 1
        return acc

    def msm(self, scalars, points):
        acc = None
        for s, P in zip(scalars, points):
            acc = self.add(acc, self.mul(P, s))
        return acc


G1 = Curve(_FqOps, G1_B, G1_GEN)
G2 = Curve(_Fq2Ops, Fq2(*G2_B), (Fq2(*G2_GEN[0]), Fq2(*G2_GEN[1])))


# ----------------------------------------------------------------------------- boundary encodings
def g1_affine_to_words(P) -> list[int]:
    """affine_t = {x, y}, 8xu32 LE each; (0,0) is infinity (affine.h / SURVEY 8b)."""
    if P is None:
        return [0] * 16
    return to_limbs(P[0]) + to_limbs(P[1])


def g2_affine_to_words(P) -> list[int]:
    if P is None:
        return [0] * 32
    x, y = P
    return to_limbs(x.c0) + to_limbs(x.c1) + to_limbs(y.c0) + to_limbs(y.c1)


def g1_projective_words_to_affine(w):
    """projective_t = {x,y,z} homogeneous; z==0 is infinity (projective.h:26-38)."""
    x, y, z = from_limbs(w[0:8]), from_limbs(w[8:16]), from_limbs(w[16:24])
    if z == 0:
        return None
    zi = finv(z, Q_MOD)
    return (x * zi % Q_MOD, y * zi % Q_MOD)


def g2_projective_words_to_affine(w):
    c = [from_limbs(w[8 * i:8 * i + 8]) for i in range(6)]
    x, y, z = Fq2(c[0], c[1]), Fq2(c[2], c[3]), Fq2(c[4], c[5])
    if z.is_zero():
        return None
    zi = z.inv()
    return (x * zi, y * zi)
