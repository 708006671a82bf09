#!/usr/bin/env python3
"""Generator for csrc/field_asm.inc.h — the inline-PTX carry chains behind the
BN254 Fr/Fq Montgomery multiplication on 8x32-bit limbs.

Why generated: there is no GPU in the build container, so the exact PTX text is
also executed by a small PTX-subset interpreter (`run_block`) in
tests/test_field_asm.py, which pins the emitted instruction sequence against
Python big-int arithmetic before it ever reaches a B200.

Scheme (replaces the reference's Karatsuba + Barrett,
/root/reference/icicle/backend/cuda/include/cuda_math.h:299-346,491-526):
word-serial Montgomery (CIOS) with the running total split into two 8-limb
accumulators X ("aligned at limb 0") and Y ("aligned at limb 1"),
T = X + Y*2^32.  Each 32x32->64 product then lands on an aligned register pair,
so every (mad.lo.cc, madc.hi.cc) pair fuses into one IMAD.WIDE.U32.X in SASS
and a whole row is a single carry chain.  After the reduction row X[0]==0; the
/2^32 is a role swap (new X = Y, new Y = X>>64) with the stray limb X[1] folded
into the next row's chain.  136 wide multiply-adds per product.
"""
from __future__ import annotations

import os
import sys

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
FIELDS = {"fr": R_MOD, "fq": Q_MOD}
N = 8
MASK = 0xFFFFFFFF


def limbs(x):
    return [(x >> (32 * i)) & MASK for i in range(N)]


def inv32(p):
    return (-pow(p, -1, 1 << 32)) % (1 << 32)


def h(v):
    return "0x%08x" % v


# ----------------------------------------------------------------------------- blocks
# A block = (operands, instructions). Operand names: X0..X7, Y0..Y7 (in/out), A0..A7, B (in).
def block_mulacc():
    """T += a*b  fused with the pending /2^32 of the previous row.
    On entry: X = previous row's Y accumulator, Y = previous row's X accumulator
    (whose limb 0 is zero, limb 1 is the stray limb, limbs 2.. become the new Y)."""
    ins = []
    ins.append(("add.cc.u32", "X0", "X0", "Y1"))
    for k, j in enumerate((1, 3, 5, 7)):
        lo, hi = 2 * k, 2 * k + 1
        c_lo = f"Y{lo + 2}" if lo + 2 < N else "0"
        c_hi = f"Y{hi + 2}" if hi + 2 < N else "0"
        ins.append(("madc.lo.cc.u32", f"Y{lo}", f"A{j}", "B", c_lo))
        ins.append(("madc.hi.cc.u32" if hi < N - 1 else "madc.hi.u32", f"Y{hi}", f"A{j}", "B", c_hi))
    for k, j in enumerate((0, 2, 4, 6)):
        lo, hi = 2 * k, 2 * k + 1
        ins.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", f"X{lo}", f"A{j}", "B", f"X{lo}"))
        ins.append(("madc.hi.cc.u32", f"X{hi}", f"A{j}", "B", f"X{hi}"))
    ins.append(("addc.u32", "Y7", "Y7", "0"))
    ops = [(f"X{i}", "+r") for i in range(N)] + [(f"Y{i}", "+r") for i in range(N)] + \
          [(f"A{i}", "r") for i in range(N)] + [("B", "r")]
    return ops, ins, []


def block_mulfirst():
    """First row: X = a_even*b, Y = a_odd*b (no accumulate, no stray)."""
    ins = []
    for k, j in enumerate((1, 3, 5, 7)):
        ins.append(("mul.lo.u32", f"Y{2 * k}", f"A{j}", "B"))
        ins.append(("mul.hi.u32", f"Y{2 * k + 1}", f"A{j}", "B"))
    for k, j in enumerate((0, 2, 4, 6)):
        ins.append(("mul.lo.u32", f"X{2 * k}", f"A{j}", "B"))
        ins.append(("mul.hi.u32", f"X{2 * k + 1}", f"A{j}", "B"))
    ops = [(f"X{i}", "=r") for i in range(N)] + [(f"Y{i}", "=r") for i in range(N)] + \
          [(f"A{i}", "r") for i in range(N)] + [("B", "r")]
    return ops, ins, []


def block_redc(p):
    """m = X0 * (-p^-1); T += m*p  => X0 == 0."""
    pl = limbs(p)
    ins = [("mul.lo.u32", "m", "X0", h(inv32(p)))]
    for k, j in enumerate((1, 3, 5, 7)):
        lo, hi = 2 * k, 2 * k + 1
        ins.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", f"Y{lo}", "m", h(pl[j]), f"Y{lo}"))
        ins.append(("madc.hi.cc.u32" if hi < N - 1 else "madc.hi.u32", f"Y{hi}", "m", h(pl[j]), f"Y{hi}"))
    for k, j in enumerate((0, 2, 4, 6)):
        lo, hi = 2 * k, 2 * k + 1
        ins.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", f"X{lo}", "m", h(pl[j]), f"X{lo}"))
        ins.append(("madc.hi.cc.u32", f"X{hi}", "m", h(pl[j]), f"X{hi}"))
    ins.append(("addc.u32", "Y7", "Y7", "0"))
    ops = [(f"X{i}", "+r") for i in range(N)] + [(f"Y{i}", "+r") for i in range(N)]
    return ops, ins, ["m"]


def block_merge():
    """R = Y + (X >> 32): the last pending /2^32.  R < 2p."""
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32")
        ins.append((op, f"R{k}", f"Y{k}", f"X{k + 1}" if k + 1 < N else "0"))
    ops = [(f"R{i}", "=r") for i in range(N)] + [(f"X{i}", "r") for i in range(N)] + [(f"Y{i}", "r") for i in range(N)]
    return ops, ins, []


def block_subp(p):
    """T = R - p, BW = 0 if R >= p else 0xffffffff."""
    pl = limbs(p)
    ins = []
    for k in range(N):
        ins.append(("sub.cc.u32" if k == 0 else "subc.cc.u32", f"T{k}", f"R{k}", h(pl[k])))
    ins.append(("subc.u32", "BW", "0", "0"))
    ops = [(f"T{i}", "=r") for i in range(N)] + [("BW", "=r")] + [(f"R{i}", "r") for i in range(N)]
    return ops, ins, []


def block_add():
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32")
        ins.append((op, f"R{k}", f"A{k}", f"C{k}"))
    ops = [(f"R{i}", "=r") for i in range(N)] + [(f"A{i}", "r") for i in range(N)] + [(f"C{i}", "r") for i in range(N)]
    return ops, ins, []


def block_sub():
    """R = A - C (mod 2^256), BW = borrow mask."""
    ins = []
    for k in range(N):
        ins.append(("sub.cc.u32" if k == 0 else "subc.cc.u32", f"R{k}", f"A{k}", f"C{k}"))
    ins.append(("subc.u32", "BW", "0", "0"))
    ops = [(f"R{i}", "=r") for i in range(N)] + [("BW", "=r")] + [(f"A{i}", "r") for i in range(N)] + \
          [(f"C{i}", "r") for i in range(N)]
    return ops, ins, []


def block_addp_masked(p):
    """R += p & MK  (MK is 0 or 0xffffffff)."""
    pl = limbs(p)
    ins = []
    for k in range(N):
        ins.append(("and.b32", f"t{k}", "MK", h(pl[k])))
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32")
        ins.append((op, f"R{k}", f"R{k}", f"t{k}"))
    ops = [(f"R{i}", "+r") for i in range(N)] + [("MK", "r")]
    return ops, ins, [f"t{k}" for k in range(N)]



# ----------------------------------------------------------------------------- wide (512-bit) products, SOS reduction
# T = X + Y*2^32 with X[0..15] aligned at limb 0 and Y[0..15] aligned at limb 1 (same even/odd pairing as above, so every
# mad.lo.cc/madc.hi.cc pair lands on an aligned register pair).  Used for lazy reduction in Fq2 (3 wide products, 2
# reductions instead of 3) and for the dedicated squaring.
W = 16


def _row_targets(i):
    """Row i adds x_j * b * 2^(32(i+j)), j = 0..7.  Returns (jlist, array, first limb index) for the two carry chains."""
    if i % 2 == 0:
        return ((0, 2, 4, 6), "X", i), ((1, 3, 5, 7), "Y", i)
    return ((1, 3, 5, 7), "X", i + 1), ((0, 2, 4, 6), "Y", i - 1)


def block_mulwide_row(i, first=False, src="A", const_limbs=None):
    """(X,Y) += src * B * 2^(32 i); src = operand array A, or the modulus limbs as immediates (reduction row)."""
    ins, used = [], set()
    for jl, arr, base in _row_targets(i):
        for k, j in enumerate(jl):
            lo, hi = base + 2 * k, base + 2 * k + 1
            a = h(const_limbs[j]) if const_limbs else f"{src}{j}"
            if first:
                ins.append(("mul.lo.u32", f"{arr}{lo}", a, "B"))
                ins.append(("mul.hi.u32", f"{arr}{hi}", a, "B"))
            else:
                ins.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", f"{arr}{lo}", a, "B", f"{arr}{lo}"))
                ins.append(("madc.hi.cc.u32", f"{arr}{hi}", a, "B", f"{arr}{hi}"))
            used.update((f"{arr}{lo}", f"{arr}{hi}"))
        nxt = base + 8
        if not first and nxt < W and const_limbs:
            # reduction row: the limbs above the chain hold live data (an addc could wrap once in 2^32 calls), so the
            # carry is counted in a separate word K[limb - 8] that redc16_final adds to the result
            kidx = (nxt if arr == "X" else nxt + 1) - 8
            ins.append(("addc.u32", f"K{kidx}", f"K{kidx}", "0"))
            used.add(f"K{kidx}")
        elif not first and nxt < W:
            # product row on a zero-initialised accumulator: the pair above the chain holds at most earlier carries,
            # one addc cannot overflow it
            ins.append(("addc.u32", f"{arr}{nxt}", f"{arr}{nxt}", "0"))
            used.add(f"{arr}{nxt}")
        elif not first:
            # chain ends at the top of the 16-limb accumulator: no carry can leave a < 2^512 value; close the chain
            last = ins.pop()
            ins.append((last[0].replace(".cc", ""),) + last[1:])
    names = sorted(used, key=lambda n: (n[0], int(n[1:])))
    ops = [(n, "=r" if first else "+r") for n in names]
    if not const_limbs:
        ops += [(f"{src}{j}", "r") for j in range(N)]
    ops += [("B", "r")]
    return ops, ins, []


def _chain8(kind, outs, xs, ys, cin=None, cout=None):
    """outs = xs (+|-) ys over 8 limbs with optional carry/borrow word in (0/1) and out (0/1)."""
    add = kind == "add"
    ins, temps = [], []
    if cin is not None:
        temps.append("tc")
        # CC := cin  (0 + 0xffffffff carries iff cin == 1;  0 - cin borrows iff cin == 1)
        ins.append(("add.cc.u32", "tc", cin, "0xffffffff") if add else ("sub.cc.u32", "tc", "0", cin))
    for k in range(8):
        first = k == 0 and cin is None
        last = k == 7 and cout is None
        base = ("add" if add else "sub") if first else ("addc" if add else "subc")
        op = base + ("" if last else ".cc") + ".u32"
        ins.append((op, outs[k], xs[k], ys[k]))
    if cout is not None:
        if add:
            ins.append(("addc.u32", cout, "0", "0"))
        else:
            ins.append(("subc.u32", cout, "0", "0"))       # 0 or 0xffffffff
            ins.append(("and.b32", cout, cout, "0x00000001"))
    return ins, temps


def block_merge16_lo():
    """T[0..7] of X + Y*2^32 and the carry word CO into limb 8."""
    ins = [("mov.u32", "T0", "X0")]
    for k in range(1, 8):
        ins.append(("add.cc.u32" if k == 1 else "addc.cc.u32", f"T{k}", f"X{k}", f"Y{k - 1}"))
    ins.append(("addc.u32", "CO", "0", "0"))
    ops = [(f"T{i}", "=r") for i in range(8)] + [("CO", "=r")] + [(f"X{i}", "r") for i in range(8)] + [(f"Y{i}", "r") for i in range(7)]
    return ops, ins, []


def block_merge16_hi():
    """T[8..15] = X[8..15] + Y[7..14] + CI (the value is < 2^512: no carry leaves limb 15)."""
    ins, temps = _chain8("add", [f"T{8 + k}" for k in range(8)], [f"X{8 + k}" for k in range(8)], [f"Y{7 + k}" for k in range(8)], cin="CI")
    ops = [(f"T{8 + i}", "=r") for i in range(8)] + [(f"X{8 + i}", "r") for i in range(8)] + [(f"Y{7 + i}", "r") for i in range(8)] + [("CI", "r")]
    return ops, ins, temps


def block_addsub8(kind, with_cin, with_cout):
    """R (+|-)= U over 8 limbs, carry/borrow words CI / CO as 0/1."""
    ins, temps = _chain8(kind, [f"R{k}" for k in range(8)], [f"R{k}" for k in range(8)], [f"U{k}" for k in range(8)],
                         cin="CI" if with_cin else None, cout="CO" if with_cout else None)
    ops = [(f"R{i}", "+r") for i in range(8)]
    if with_cout:
        ops.append(("CO", "=r"))
    ops += [(f"U{i}", "r") for i in range(8)]
    if with_cin:
        ops.append(("CI", "r"))
    return ops, ins, temps


def block_addconst8(c8, with_cin, with_cout):
    """R += 8-limb constant, carry words as above."""
    cl = [(c8 >> (32 * i)) & MASK for i in range(8)]
    ins, temps = _chain8("add", [f"R{k}" for k in range(8)], [f"R{k}" for k in range(8)], [h(x) for x in cl],
                         cin="CI" if with_cin else None, cout="CO" if with_cout else None)
    ops = [(f"R{i}", "+r") for i in range(8)]
    if with_cout:
        ops.append(("CO", "=r"))
    if with_cin:
        ops.append(("CI", "r"))
    return ops, ins, temps


def block_redc16_step(i, p):
    """One SOS reduction step on (X, Y, C): with all limbs below i already zero,
    m = (X_i + Y_{i-1} + C mod 2^32) * (-p^-1);  (X,Y) += m * p * 2^(32 i);  C = carry of limb i into limb i+1."""
    pl = limbs(p)
    yi = f"Y{i - 1}" if i > 0 else "0"
    ins = [("add.u32", "t", f"X{i}", yi), ("add.u32", "t", "t", "C"), ("mul.lo.u32", "B", "t", h(inv32(p)))]
    _, row, _ = block_mulwide_row(i, const_limbs=pl)
    ins += row
    ins += [("add.cc.u32", "t", f"X{i}", yi), ("addc.u32", "c1", "0", "0"), ("add.cc.u32", "t", "t", "C"),
            ("addc.u32", "C", "c1", "0")]
    used = set()
    for ins_ in ins:
        for x in ins_[1:]:
            if x and x[0] in "XYK" and x[1:].isdigit():
                used.add(x)
    names = sorted(used, key=lambda n: (n[0], int(n[1:])))
    ops = [(n, "+r") for n in names] + [("C", "+r")]
    return ops, ins, ["t", "B", "c1"]


def block_redc16_final():
    """R = X[8..15] + Y[7..14]  (the value / 2^256, still missing the carry word C; < 2p so nothing overflows)."""
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32")
        ins.append((op, f"R{k}", f"X{8 + k}", f"Y{7 + k}"))
    ops = [(f"R{i}", "=r") for i in range(N)] + [(f"X{8 + i}", "r") for i in range(N)] + [(f"Y{7 + i}", "r") for i in range(N)]
    return ops, ins, []


def block_addk():
    """R += K[0..7] (the carry words collected by the reduction rows)."""
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32")
        ins.append((op, f"R{k}", f"R{k}", f"K{k}"))
    ops = [(f"R{i}", "+r") for i in range(N)] + [(f"K{i}", "r") for i in range(N)]
    return ops, ins, []


def block_addword():
    """R += C (one word)."""
    ins = [("add.cc.u32", "R0", "R0", "C")]
    for k in range(1, N):
        ins.append(("addc.cc.u32" if k < N - 1 else "addc.u32", f"R{k}", f"R{k}", "0"))
    ops = [(f"R{i}", "+r") for i in range(N)] + [("C", "r")]
    return ops, ins, []


def block_sqr_cross_row(i):
    """(X,Y) += sum_{j>i} a_j * a_i * 2^(32(i+j)) : the cross products of a square, row i (i = 0..6)."""
    ins, used = [], set()
    groups = {}
    for j in range(i + 1, N):
        pth = i + j
        arr, lo = ("X", pth) if pth % 2 == 0 else ("Y", pth - 1)
        groups.setdefault(arr, []).append((j, lo))
    for arr, lst in groups.items():
        for k, (j, lo) in enumerate(lst):
            ins.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", f"{arr}{lo}", f"A{j}", f"A{i}", f"{arr}{lo}"))
            ins.append(("madc.hi.cc.u32", f"{arr}{lo + 1}", f"A{j}", f"A{i}", f"{arr}{lo + 1}"))
            used.update((f"{arr}{lo}", f"{arr}{lo + 1}"))
        nxt = lst[-1][1] + 2
        if nxt < W:
            ins.append(("addc.u32", f"{arr}{nxt}", f"{arr}{nxt}", "0"))
            used.add(f"{arr}{nxt}")
        else:
            last = ins.pop()
            ins.append((last[0].replace(".cc", ""),) + last[1:])
    names = sorted(used, key=lambda n: (n[0], int(n[1:])))
    ops = [(n, "+r") for n in names] + [(f"A{j}", "r") for j in range(N)]
    return ops, ins, []


def block_double16(arr):
    """arr[0..15] *= 2 (no overflow: the doubled cross sum of a square is < 2^512)."""
    ins = []
    for k in range(W):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < W - 1 else "addc.u32")
        ins.append((op, f"{arr}{k}", f"{arr}{k}", f"{arr}{k}"))
    ops = [(f"{arr}{i}", "+r") for i in range(W)]
    return ops, ins, []


def block_sqr_diag():
    """X += sum_i a_i^2 * 2^(64 i): the squares sit exactly on the even-aligned pairs."""
    ins = []
    for i in range(N):
        ins.append(("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32", f"X{2 * i}", f"A{i}", f"A{i}", f"X{2 * i}"))
        ins.append(("madc.hi.cc.u32" if i < N - 1 else "madc.hi.u32", f"X{2 * i + 1}", f"A{i}", f"A{i}", f"X{2 * i + 1}"))
    ops = [(f"X{i}", "+r") for i in range(W)] + [(f"A{j}", "r") for j in range(N)]
    return ops, ins, []


# ----------------------------------------------------------------------------- interpreter
def run_block(block, env, allow_wrap=False):
    """Execute a block on a dict name->u32. Mirrors PTX semantics of the used subset.
    A carry out of a non-.cc addc is an error unless allow_wrap (mod-2^256 arithmetic intended)."""
    ops, ins, temps = block
    cc = 0
    reg = dict(env)

    def val(x):
        if x.startswith("0x"):
            return int(x, 16)
        if x == "0":
            return 0
        return reg[x]

    for i in ins:
        op, d, srcs = i[0], i[1], [val(s) for s in i[2:]]
        parts = op.split(".")
        base = parts[0]
        sets_cc = "cc" in parts[1:]
        uses_cc = base.endswith("c") and base in ("addc", "subc", "madc")
        cin = cc if uses_cc else 0
        if base in ("add", "addc"):
            full = srcs[0] + srcs[1] + cin
            res, cout = full & MASK, full >> 32
        elif base in ("sub", "subc"):
            full = srcs[0] - srcs[1] - cin
            res, cout = full & MASK, 1 if full < 0 else 0
        elif base in ("mad", "madc"):
            prod = srcs[0] * srcs[1]
            part = (prod & MASK) if "lo" in parts else (prod >> 32)
            full = part + srcs[2] + cin
            res, cout = full & MASK, full >> 32
        elif base == "mul":
            prod = srcs[0] * srcs[1]
            res, cout = ((prod & MASK) if "lo" in parts else (prod >> 32)), cc
        elif base == "and":
            res, cout = srcs[0] & srcs[1], cc
        elif base == "mov":
            res, cout = srcs[0], cc
        else:
            raise ValueError(op)
        if sets_cc:
            cc = cout
        elif uses_cc:
            if cout and base != "subc" and not allow_wrap:
                raise OverflowError(f"carry lost at {i}")
            cc = cc  # PTX leaves CC unchanged without .cc
        reg[d] = res
    return reg


def emulate_mont_mul(a, b, p):
    """Run the emitted blocks exactly as field.cuh sequences them."""
    al, bl = limbs(a), limbs(b)
    X = {f"X{i}": 0 for i in range(N)}
    Y = {f"Y{i}": 0 for i in range(N)}
    env = {**X, **Y, **{f"A{i}": al[i] for i in range(N)}}
    names = (("X", "Y"), ("Y", "X"))
    for i in range(N):
        cur, oth = names[i & 1]
        # rename so that block's "X" is `cur`
        view = {f"X{k}": env[f"{cur}{k}"] for k in range(N)}
        view.update({f"Y{k}": env[f"{oth}{k}"] for k in range(N)})
        view.update({f"A{k}": al[k] for k in range(N)})
        view["B"] = bl[i]
        out = run_block(block_mulfirst() if i == 0 else block_mulacc(), view)
        out = run_block(block_redc(p), out)
        for k in range(N):
            env[f"{cur}{k}"] = out[f"X{k}"]
            env[f"{oth}{k}"] = out[f"Y{k}"]
        assert out["X0"] == 0
    # after 8 rows the last row's X is names[1][0] == "Y" array
    cur, oth = names[(N - 1) & 1]
    view = {f"X{k}": env[f"{cur}{k}"] for k in range(N)}
    view.update({f"Y{k}": env[f"{oth}{k}"] for k in range(N)})
    out = run_block(block_merge(), view)
    out = run_block(block_subp(p), out)
    r = sum(out[f"R{k}"] << (32 * k) for k in range(N))
    t = sum(out[f"T{k}"] << (32 * k) for k in range(N))
    return r if out["BW"] else t


# ----------------------------------------------------------------------------- C++ emission
def emit_asm(block, argmap):
    """argmap: operand name -> C++ expression."""
    ops, ins, temps = block
    idx = {name: i for i, (name, _) in enumerate(ops)}

    def ref(x):
        if x in idx:
            return f"%{idx[x]}"
        return x  # literal or temp

    lines = []
    if temps:
        lines.append("{ .reg .u32 " + ", ".join(temps) + ";")
    for i in ins:
        lines.append(f"{i[0]} {', '.join(ref(x) for x in i[1:])};")
    if temps:
        lines.append("}")
    body = "\n".join(f'      "{l}\\n\\t"' for l in lines)
    outs = ", ".join(f'"{"=&r" if c == "=r" else c}"({argmap[n]})' for n, c in ops if c in ("+r", "=r"))
    insn = ", ".join(f'"{c}"({argmap[n]})' for n, c in ops if c == "r")
    return f"  asm(\n{body}\n      : {outs}\n      : {insn});\n"


def arr(name, prefix):
    return {f"{prefix}{i}": f"{name}[{i}]" for i in range(N)}


def generate():
    o = []
    o.append("// GENERATED by tools/gen_field_asm.py — do not edit. See that file for the scheme.\n")
    o.append("// Every asm statement is a self-contained carry chain (CC never crosses statements).\n")
    o.append("#pragma once\n#include <cstdint>\n\nnamespace b200 { namespace ptx {\n\n")
    o.append("#ifdef __CUDA_ARCH__\n")
    am = {**arr("X", "X"), **arr("Y", "Y"), **arr("a", "A"), "B": "b"}
    o.append("__device__ __forceinline__ void mul_first(uint32_t (&X)[8], uint32_t (&Y)[8], const uint32_t (&a)[8], uint32_t b) {\n")
    o.append(emit_asm(block_mulfirst(), am))
    o.append("}\n\n")
    o.append("__device__ __forceinline__ void mul_acc(uint32_t (&X)[8], uint32_t (&Y)[8], const uint32_t (&a)[8], uint32_t b) {\n")
    o.append(emit_asm(block_mulacc(), am))
    o.append("}\n\n")
    o.append("__device__ __forceinline__ void merge(uint32_t (&R)[8], const uint32_t (&X)[8], const uint32_t (&Y)[8]) {\n")
    o.append(emit_asm(block_merge(), {**arr("R", "R"), **arr("X", "X"), **arr("Y", "Y")}))
    o.append("}\n\n")
    o.append("__device__ __forceinline__ void add8(uint32_t (&R)[8], const uint32_t (&a)[8], const uint32_t (&c)[8]) {\n")
    o.append(emit_asm(block_add(), {**arr("R", "R"), **arr("a", "A"), **arr("c", "C")}))
    o.append("}\n\n")
    o.append("__device__ __forceinline__ uint32_t sub8(uint32_t (&R)[8], const uint32_t (&a)[8], const uint32_t (&c)[8]) {\n  uint32_t bw;\n")
    o.append(emit_asm(block_sub(), {**arr("R", "R"), **arr("a", "A"), **arr("c", "C"), "BW": "bw"}))
    o.append("  return bw;\n}\n\n")
    for name, p in FIELDS.items():
        o.append(f"__device__ __forceinline__ void redc_{name}(uint32_t (&X)[8], uint32_t (&Y)[8]) {{\n")
        o.append(emit_asm(block_redc(p), {**arr("X", "X"), **arr("Y", "Y")}))
        o.append("}\n\n")
        o.append(f"__device__ __forceinline__ uint32_t subp_{name}(uint32_t (&T)[8], const uint32_t (&R)[8]) {{\n  uint32_t bw;\n")
        o.append(emit_asm(block_subp(p), {**arr("T", "T"), **arr("R", "R"), "BW": "bw"}))
        o.append("  return bw;\n}\n\n")
        o.append(f"__device__ __forceinline__ void addp_masked_{name}(uint32_t (&R)[8], uint32_t mk) {{\n")
        o.append(emit_asm(block_addp_masked(p), {**arr("R", "R"), "MK": "mk"}))
        o.append("}\n\n")
    # ---- wide products / SOS reduction / squaring
    XY = {**{f"X{i}": f"X[{i}]" for i in range(W)}, **{f"Y{i}": f"Y[{i}]" for i in range(W)}}
    am16 = {**XY, **arr("a", "A"), "B": "b"}
    o.append("__device__ __forceinline__ void mulwide_row0(uint32_t (&X)[16], uint32_t (&Y)[16], const uint32_t (&a)[8], uint32_t b) {\n")
    o.append(emit_asm(block_mulwide_row(0, first=True), am16))
    o.append("}\n\n")
    for i in range(1, N):
        o.append(f"__device__ __forceinline__ void mulwide_row{i}(uint32_t (&X)[16], uint32_t (&Y)[16], const uint32_t (&a)[8], uint32_t b) {{\n")
        o.append(emit_asm(block_mulwide_row(i), am16))
        o.append("}\n\n")
    for i in range(N - 1):
        o.append(f"__device__ __forceinline__ void sqr_cross_row{i}(uint32_t (&X)[16], uint32_t (&Y)[16], const uint32_t (&a)[8]) {{\n")
        o.append(emit_asm(block_sqr_cross_row(i), {**XY, **arr("a", "A")}))
        o.append("}\n\n")
    o.append("__device__ __forceinline__ void double16(uint32_t (&X)[16]) {\n")
    o.append(emit_asm(block_double16("X"), XY))
    o.append("}\n\n")
    o.append("__device__ __forceinline__ void sqr_diag(uint32_t (&X)[16], const uint32_t (&a)[8]) {\n")
    o.append(emit_asm(block_sqr_diag(), {**XY, **arr("a", "A")}))
    o.append("}\n\n")
    T16 = {f"T{i}": f"T[{i}]" for i in range(W)}
    o.append("__device__ __forceinline__ void merge16(uint32_t (&T)[16], const uint32_t (&X)[16], const uint32_t (&Y)[16]) {\n  uint32_t co;\n")
    o.append(emit_asm(block_merge16_lo(), {**T16, **XY, "CO": "co"}))
    o.append(emit_asm(block_merge16_hi(), {**T16, **XY, "CI": "co"}))
    o.append("}\n\n")
    RU = {**arr("R", "R"), **arr("U", "U")}
    for kind in ("add", "sub"):
        o.append(f"// T {'+' if kind == 'add' else '-'}= U over 16 limbs (no overflow / no underflow by the caller's bounds)\n")
        o.append(f"__device__ __forceinline__ void {kind}16(uint32_t (&T)[16], const uint32_t (&U)[16]) {{\n  uint32_t co;\n")
        lo = {**{f"R{i}": f"T[{i}]" for i in range(8)}, **{f"U{i}": f"U[{i}]" for i in range(8)}, "CO": "co"}
        hi = {**{f"R{i}": f"T[{8 + i}]" for i in range(8)}, **{f"U{i}": f"U[{8 + i}]" for i in range(8)}, "CI": "co"}
        o.append(emit_asm(block_addsub8(kind, False, True), lo))
        o.append(emit_asm(block_addsub8(kind, True, False), hi))
        o.append("}\n\n")
    for name, p in FIELDS.items():
        for i in range(N):
            o.append(f"__device__ __forceinline__ void redc16_step{i}_{name}(uint32_t (&X)[16], uint32_t (&Y)[16], uint32_t (&K)[8], uint32_t& c) {{\n")
            o.append(emit_asm(block_redc16_step(i, p), {**XY, **{f"K{j}": f"K[{j}]" for j in range(8)}, "C": "c"}))
            o.append("}\n\n")
        p2 = p * p
        o.append(f"// T += p^2 (16 limbs)\n__device__ __forceinline__ void addp2_{name}(uint32_t (&T)[16]) {{\n  uint32_t co;\n")
        o.append(emit_asm(block_addconst8(p2 & ((1 << 256) - 1), False, True), {**{f"R{i}": f"T[{i}]" for i in range(8)}, "CO": "co"}))
        o.append(emit_asm(block_addconst8(p2 >> 256, True, False), {**{f"R{i}": f"T[{8 + i}]" for i in range(8)}, "CI": "co"}))
        o.append("}\n\n")
    o.append("__device__ __forceinline__ void redc16_final(uint32_t (&R)[8], const uint32_t (&X)[16], const uint32_t (&Y)[16], const uint32_t (&K)[8], uint32_t c) {\n")
    o.append(emit_asm(block_redc16_final(), {**arr("R", "R"), **XY}))
    o.append(emit_asm(block_addk(), {**arr("R", "R"), **{f"K{j}": f"K[{j}]" for j in range(8)}}))
    o.append(emit_asm(block_addword(), {**arr("R", "R"), "C": "c"}))
    o.append("}\n\n")
    o.append("#endif // __CUDA_ARCH__\n\n}} // namespace b200::ptx\n")
    return "".join(o)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "field_asm.inc.h")
    if len(sys.argv) > 1:
        out = sys.argv[1]
    with open(out, "w") as f:
        f.write(generate())
    print("wrote", out)
