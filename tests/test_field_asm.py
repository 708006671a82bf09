"""The generated PTX carry chains (csrc/field_asm.inc.h), executed by the generator's PTX-subset
interpreter, against Python big-int Montgomery arithmetic. Pins the device multiplier's instruction
sequence before it reaches a GPU."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gen_field_asm", os.path.join(ROOT, "icicle-snark_b200", "tools", "gen_field_asm.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)

RINV = {name: pow(1 << 256, -1, p) for name, p in G.FIELDS.items()}


def cases(p, rnd):
    edge = [0, 1, 2, p - 1, p - 2, (1 << 254) % p, (1 << 256) % p, (p - 1) // 2, 0xFFFFFFFF, (1 << 224) - 1]
    for a in edge:
        for b in edge:
            yield a, b
    for _ in range(300):
        yield rnd.randrange(p), rnd.randrange(p)


def test_montgomery_product_matches_bigint():
    rnd = random.Random(7)
    for name, p in G.FIELDS.items():
        for a, b in cases(p, rnd):
            assert G.emulate_mont_mul(a, b, p) == a * b * RINV[name] % p, (name, hex(a), hex(b))


def test_generated_header_is_current():
    path = os.path.join(ROOT, "icicle-snark_b200", "csrc", "field_asm.inc.h")
    assert open(path).read() == G.generate(), "run icicle-snark_b200/tools/gen_field_asm.py"


def test_add_sub_blocks():
    rnd = random.Random(9)
    for name, p in G.FIELDS.items():
        for _ in range(200):
            a, b = rnd.randrange(p), rnd.randrange(p)
            env = {f"A{i}": l for i, l in enumerate(G.limbs(a))}
            env.update({f"C{i}": l for i, l in enumerate(G.limbs(b))})
            out = G.run_block(G.block_sub(), env)
            r = {f"R{i}": out[f"R{i}"] for i in range(8)}
            r["MK"] = out["BW"]
            out2 = G.run_block(G.block_addp_masked(p), r, allow_wrap=True)
            got = sum(out2[f"R{i}"] << (32 * i) for i in range(8))
            assert got == (a - b) % p


# ---- wide products, SOS reduction, squaring (lazy reduction in Fq2, dedicated squaring) -------------------------
M32 = (1 << 32) - 1


def _xy_value(env):
    return sum(env[f"X{k}"] << (32 * k) for k in range(16)) + sum(env[f"Y{k}"] << (32 * (k + 1)) for k in range(16))


def _mulwide(a, b):
    al, bl = G.limbs(a), G.limbs(b)
    env = {f"A{i}": al[i] for i in range(8)}
    env.update({f"{x}{k}": 0 for x in "XY" for k in range(16)})
    for i in range(8):
        env["B"] = bl[i]
        out = G.run_block(G.block_mulwide_row(i, first=(i == 0)), env)
        env.update({k: v for k, v in out.items() if k[0] in "XY"})
    return env


def _sqrwide(a):
    env = {f"A{i}": l for i, l in enumerate(G.limbs(a))}
    env.update({f"{x}{k}": 0 for x in "XY" for k in range(16)})
    for i in range(7):
        out = G.run_block(G.block_sqr_cross_row(i), env)
        env.update({k: v for k, v in out.items() if k[0] in "XY"})
    for arr in "XY":
        out = G.run_block(G.block_double16(arr), env)
        env.update({k: v for k, v in out.items() if k[0] in "XY"})
    out = G.run_block(G.block_sqr_diag(), env)
    env.update({k: v for k, v in out.items() if k[0] in "XY"})
    return env


def _redc16(env, p):
    env = dict(env, C=0, **{f"K{j}": 0 for j in range(9)})
    for i in range(8):
        out = G.run_block(G.block_redc16_step(i, p), env)
        env.update({k: v for k, v in out.items() if k[0] in "XYCK"})
    assert env["K8"] == 0
    out = G.run_block(G.block_redc16_final(), env)
    r = {f"R{k}": out[f"R{k}"] for k in range(8)}
    r.update({f"K{k}": env[f"K{k}"] for k in range(8)})
    out = G.run_block(G.block_addk(), r)
    r = {f"R{k}": out[f"R{k}"] for k in range(8)}
    r["C"] = env["C"]
    out = G.run_block(G.block_addword(), r)
    return sum(out[f"R{k}"] << (32 * k) for k in range(8))


def test_wide_product_and_square_are_exact():
    rnd = random.Random(11)
    edge = [0, 1, (1 << 256) - 1, (1 << 255), M32, M32 << 224]
    for a in edge + [rnd.randrange(1 << 256) for _ in range(150)]:
        for b in edge[:3] + [rnd.randrange(1 << 256)]:
            env = _mulwide(a, b)
            assert _xy_value(env) == a * b
            lo = G.run_block(G.block_merge16_lo(), env)
            hi = G.run_block(G.block_merge16_hi(), dict(env, CI=lo["CO"]))
            t = sum(lo[f"T{k}"] << (32 * k) for k in range(8)) + sum(hi[f"T{k}"] << (32 * k) for k in range(8, 16))
            assert t == a * b
        assert _xy_value(_sqrwide(a)) == a * a


def test_sos_reduction_including_saturated_upper_limbs():
    rnd = random.Random(12)
    for name, p in G.FIELDS.items():
        rinv = RINV[name]
        cases = [(p << 256) - 1, 0, p, (p - 1) << 256]
        cases += [rnd.randrange(p << 256) for _ in range(300)]
        # limbs just above a reduction row saturated: the case a plain addc into live data would get wrong
        cases += [(rnd.randrange(p << 256) | (M32 << (32 * rnd.randrange(8, 15)))) % (p << 256) for _ in range(300)]
        for t in cases:
            y = rnd.randrange(min(t >> 32, 1 << 479) + 1) if rnd.random() < 0.5 else 0
            x = t - (y << 32)
            env = {f"X{k}": (x >> (32 * k)) & M32 for k in range(16)}
            env.update({f"Y{k}": (y >> (32 * k)) & M32 for k in range(16)})
            r = _redc16(env, p)
            assert r < 2 * p and r % p == t * rinv % p


def test_lazy_fq2_product_bounds_and_value():
    # c0 = (a0 b0 + p^2 - a1 b1)/R, c1 = ((a0+a1)(b0+b1) - a0 b0 - a1 b1)/R, both inputs of redc16 < p * 2^256
    rnd = random.Random(13)
    p = G.Q_MOD
    rinv = RINV["fq"]
    for _ in range(200):
        a0, a1, b0, b1 = (rnd.choice([0, 1, p - 1, rnd.randrange(p)]) for _ in range(4))
        t0, t1, t2 = a0 * b0, a1 * b1, (a0 + a1) * (b0 + b1)
        c0w, c1w = t0 + p * p - t1, t2 - t0 - t1
        assert 0 <= c0w < p << 256 and 0 <= c1w < p << 256
        for w, want in ((c0w, (a0 * b0 - a1 * b1) * rinv % p), (c1w, (a0 * b1 + a1 * b0) * rinv % p)):
            env = {f"X{k}": (w >> (32 * k)) & M32 for k in range(16)}
            env.update({f"Y{k}": 0 for k in range(16)})
            assert _redc16(env, p) % p == want
