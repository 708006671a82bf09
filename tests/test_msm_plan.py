"""CPU: host logic of the Pippenger plan (csrc/msm_g1.cu make_msm_plan) through b200_msm_plan_info - window count,
bucket sets after precompute folding, the signed-digit recoding constant - checked against a big-int restatement of the
digit kernel's arithmetic (csrc/msm_sort.cu msm_digits_kernel)."""
import ctypes as C
import random

import numpy as np
import pytest

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def plan(lib, n, c=0, bitsize=254, factor=1, g2=False):
    out = (C.c_int32 * 8)()
    h = (C.c_uint32 * 9)()
    err = lib.dll.b200_msm_plan_info(n, c, bitsize, factor, int(g2), out, h)
    assert err == 0, err
    keys = ("c", "windows", "factor", "sets", "bpw", "nbuckets", "item_cap", "n")
    d = dict(zip(keys, list(out)))
    d["H"] = int.from_bytes(np.array(list(h), dtype=np.uint32).tobytes(), "little")
    return d


def digits(s, p):
    """the kernel's recoding: u_w = ((s + H) >> c w) mod 2^c, d_w = u_w - 2^(c-1) in [-2^(c-1), 2^(c-1))"""
    t = s + p["H"]
    half, mask = 1 << (p["c"] - 1), (1 << p["c"]) - 1
    ds = [((t >> (w * p["c"])) & mask) - half for w in range(p["windows"])]
    assert t >> (p["c"] * p["windows"]) == 0, "the windows must cover s + H"
    return ds


@pytest.mark.parametrize("c", list(range(2, 23)))
@pytest.mark.parametrize("bitsize", [254, 64, 10])
def test_recoding_reconstructs_every_scalar(lib, c, bitsize):
    p = plan(lib, 1 << 16, c=c, bitsize=bitsize)
    assert p["c"] == c and p["bpw"] == 1 << (c - 1) and p["windows"] * c >= bitsize + 1
    assert p["H"] == sum((1 << (c - 1)) << (c * w) for w in range(p["windows"]))
    rnd = random.Random(c * 1000 + bitsize)
    top = min(R, 1 << bitsize)
    for s in [0, 1, top - 1, (1 << (bitsize - 1)) % top] + [rnd.randrange(top) for _ in range(200)]:
        ds = digits(s, p)
        assert sum(d << (c * w) for w, d in enumerate(ds)) == s
        # bucket keys are |d| - 1 in [0, bpw): a digit of magnitude 2^(c-1) (only -2^(c-1) occurs) is the last bucket
        assert all(-p["bpw"] <= d < p["bpw"] for d in ds)


def test_precompute_folds_windows_into_bucket_sets(lib):
    for factor in (1, 2, 4, 8, 13, 16, 32):
        for n in (1 << 12, 100_000, 400_000, 1 << 20, 3_200_002, 1 << 22):
            p = plan(lib, n, factor=factor)
            assert 2 <= p["c"] <= 22 and p["factor"] == min(factor, p["windows"])
            assert p["sets"] == -(-p["windows"] // p["factor"]) and p["nbuckets"] == p["sets"] * p["bpw"]
            assert p["windows"] == -(-(254 + 2) // p["c"])


def test_heuristic_matches_the_measured_sweet_spots(lib):
    """Values swept on the B200: with the prover's 16 tables, c = 20 from 2^20 points and c = 17
    below; either way all windows share one bucket set, and the top window never holds just 1..8 scalar bits."""
    for n, c in ((100_000, 17), (200_000, 17), (400_000, 17), (524_288, 17), (800_000, 20), (1 << 20, 20),
                 (1_600_001, 20), (3_200_002, 20), (1 << 22, 20)):
        p = plan(lib, n, factor=16)
        assert (p["c"], p["sets"]) == (c, 1), (n, p)
        top_bits = 254 - p["c"] * (p["windows"] - 1)
        assert top_bits <= 0 or top_bits > 8
    # no tables: about log2(n) - 5, G2 one less (its adds cost 3x, so the reduction must stay small)
    assert plan(lib, 1 << 22)["c"] == 17 and plan(lib, 1 << 22, g2=True)["c"] == 16
    assert plan(lib, 1 << 26)["c"] == 20  # 21 would leave a 2-bit top window
    assert plan(lib, 1)["c"] == 2 and plan(lib, 1)["windows"] == 128


def test_plan_info_argument_checks(lib):
    out = (C.c_int32 * 8)()
    assert lib.dll.b200_msm_plan_info(0, 0, 254, 1, 0, out, None) == 11
    assert lib.dll.b200_msm_plan_info(16, 23, 254, 1, 0, out, None) == 11
    assert lib.dll.b200_msm_plan_info(16, 0, 255, 1, 0, out, None) == 11
    assert lib.dll.b200_msm_plan_info(16, 0, 254, 1, 0, None, None) == 3
