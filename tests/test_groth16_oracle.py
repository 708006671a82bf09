"""CPU: the oracle pipeline (reference C++ driven through oracle/groth16_ref.py) against the committed
golden fixtures, the file-format restatement, and the pairing check."""
import os

import numpy as np
import pytest

import icicle_snark_b200 as pkg
from oracle import groth16_ref as G
from tools import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXED_R = 0x1d2c3b4a5968778695a4b3c2d1e0f00112233445566778899aabbccddeeff001 % synth.R
FIXED_S = 0x0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcdef % synth.R


def load(n):
    base = os.path.join(GOLD, f"complex_{n}")
    vkz = np.load(base + ".vk.npz")
    vk = {k: vkz[k] for k in ("alpha1", "beta2", "gamma2", "delta2", "ic")}
    vk["n_public"] = int(vkz["n_public"])
    return (open(base + ".zkey", "rb").read(), open(base + ".wtns", "rb").read(), vk,
            open(base + ".proof_r1s1.json").read(), open(base + ".proof_rs.json").read(), open(base + ".public.json").read())


@pytest.mark.parametrize("n", [6, 100])
def test_oracle_reproduces_golden_and_verifies(ref, n):
    zkey, wtns, vk, gold11, goldrs, goldpub = load(n)
    proof, public = G.prove(ref, pkg.bindings, zkey, wtns, 1, 1)
    assert G.proof_json(proof) == gold11 and G.public_json(public) == goldpub
    assert G.verify(ref, proof, public, vk)
    proof2, _ = G.prove(ref, pkg.bindings, zkey, wtns, FIXED_R, FIXED_S)
    assert G.proof_json(proof2) == goldrs and G.verify(ref, proof2, public, vk)
    # a wrong public input or a swapped proof element must not verify
    assert not G.verify(ref, proof, [public[0] + 1], vk)
    bad = dict(proof, pi_c=proof2["pi_c"])
    assert not G.verify(ref, bad, public, vk)


def test_synth_is_deterministic_and_matches_committed_files(ref):
    zkey, wtns, vk, *_ = load(6)
    z2, w2, vk2 = synth.make_complex_circuit(ref, 6)
    assert z2 == zkey and w2 == wtns and np.array_equal(vk2["ic"], vk["ic"])


def test_zkey_wtns_format_restatement():
    zkey, wtns, *_ = load(6)
    z = G.parse_zkey(zkey)
    assert (z["n_vars"], z["n_public"], z["domain_size"], z["power"]) == (8, 1, 8, 3)
    assert len(z["coef"]) == 2 * 6 + 2 and z["A"].shape == (8, 16) and z["B2"].shape == (8, 32)
    assert z["C"].shape == (6, 16) and z["H"].shape == (8, 16)
    assert not z["B1"][0].any() and not z["B1"][1].any()  # v_0 = v_1 = 0 -> points at infinity (0,0)
    w = G.parse_wtns(wtns)
    assert w["n_witness"] == 8 and w["q"] == G.FR_BYTES
    vals = [int.from_bytes(x.tobytes(), "little") for x in w["w"]]
    assert vals[0] == 1 and vals[2] == 3 and vals[3] == 9 and vals[1] == pow(3, 2 ** 6, synth.R)
    with pytest.raises(ValueError):
        G.parse_binfile(b"wtns" + zkey[4:], b"zkey")
    with pytest.raises(ValueError):
        G.prove(None, pkg.bindings, zkey, wtns[:-32 * 2] , 1, 1, cache=type("C", (), {"z": dict(z, n_vars=9)})())


def test_json_layout_is_serde_pretty():
    _, _, _, gold11, _, goldpub = load(6)
    lines = gold11.split("\n")
    assert lines[0] == "{" and lines[1] == '  "curve": "bn128",' and lines[2] == '  "pi_a": [' and lines[-1] == "}"
    assert lines[-2] == '  "protocol": "groth16"' and not gold11.endswith("\n")
    keys = [l.split('"')[1] for l in lines if l.startswith('  "')]
    assert keys == sorted(keys) == ["curve", "pi_a", "pi_b", "pi_c", "protocol"]
    assert goldpub.startswith('[\n  "') and goldpub.endswith('"\n]')
