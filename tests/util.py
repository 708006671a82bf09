"""Test helpers: conversions between Python ints and the 8xu32 LE limb arrays of the C ABI, seeded inputs."""
import numpy as np

from oracle import bn254_py as O

R = O.R_MOD
Q = O.Q_MOD


def to_words(x, n=8):
    return np.array([(int(x) >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def from_words(w):
    v = 0
    for i, l in enumerate(np.asarray(w).reshape(-1)):
        v |= int(l) << (32 * i)
    return v


def ints_to_array(xs):
    """list of ints -> (n, 8) u32"""
    b = b"".join(int(x).to_bytes(32, "little") for x in xs)
    return np.frombuffer(b, dtype=np.uint32).reshape(-1, 8).copy()


def array_to_ints(a):
    raw = np.ascontiguousarray(a, dtype=np.uint32).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def rand_scalars(rng, n, modulus=R):
    """uniform in [0, modulus) from 256-bit draws"""
    raw = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    vals = [v % modulus for v in array_to_ints(raw)]
    return ints_to_array(vals), vals


def g1_points_multiples(n, start=1):
    """P_i = (start+i)*G by repeated addition (distinct, valid, known dlog) -> (n,16) u32 + affine tuples"""
    cv = O.G1
    pts, cur = [], cv.mul(O.G1_GEN, start)
    for _ in range(n):
        pts.append(cur)
        cur = cv.add(cur, O.G1_GEN)
    arr = np.array([O.g1_affine_to_words(p) for p in pts], dtype=np.uint32)
    return arr, pts


def g2_points_multiples(n, start=1):
    cv = O.G2
    pts, cur = [], cv.mul(cv.gen, start)
    for _ in range(n):
        pts.append(cur)
        cur = cv.add(cur, cv.gen)
    arr = np.array([O.g2_affine_to_words(p) for p in pts], dtype=np.uint32)
    return arr, pts
