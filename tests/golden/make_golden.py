"""Regenerates the committed Groth16 fixtures with the REFERENCE's own CPU library (oracle/_ref):
  complex_{n}.zkey / .wtns     synthetic ComplexCircuit(n,n) artefacts (tools/synth.py, seeded toxic waste)
  complex_{n}.vk.npz           verification key (standard-form affine words)
  complex_{n}.vk.json          the same key as snarkjs's verification_key.json (what `verify --vk` reads)
  complex_{n}.proof_r1s1.json  proof.json with r = s = 1 (the reference's `no-randomness` feature)
  complex_{n}.proof_rs.json    proof.json with the fixed (r, s) below
  complex_{n}.public.json
Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401  (registers the package)
import icicle_snark_b200 as pkg
from oracle import groth16_ref as G
from oracle import ref_cpu
from tools import synth

FIXED_R = 0x1d2c3b4a5968778695a4b3c2d1e0f00112233445566778899aabbccddeeff001 % synth.R
FIXED_S = 0x0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcdef % synth.R
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    ref = ref_cpu.ref()
    for n in (6, 100):
        zkey, wtns, vk = synth.make_complex_circuit(ref, n)
        base = os.path.join(HERE, f"complex_{n}")
        open(base + ".zkey", "wb").write(zkey)
        open(base + ".wtns", "wb").write(wtns)
        np.savez(base + ".vk.npz", **{k: v for k, v in vk.items() if k != "n_public"}, n_public=vk["n_public"])
        open(base + ".vk.json", "w").write(synth.vk_json(vk))
        for tag, (r, s) in (("r1s1", (1, 1)), ("rs", (FIXED_R, FIXED_S))):
            proof, public = G.prove(ref, pkg.bindings, zkey, wtns, r, s)
            assert G.verify(ref, proof, public, vk)
            open(base + f".proof_{tag}.json", "w").write(G.proof_json(proof))
        open(base + ".public.json", "w").write(G.public_json(public))
        print("wrote", base)
