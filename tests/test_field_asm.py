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
