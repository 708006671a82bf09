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
    o.append("#endif // __CUDA_ARCH__\n\n}} // namespace b200::ptx\n")
    return "".join(o)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "field_asm.inc.h")
    if len(sys.argv) > 1:
        out = sys.argv[1]
    with open(out, "w") as f:
        f.write(generate())
    print("wrote", out)
