"""Host-side logic of the fused path that needs no GPU: the C proof.json writer against the golden files, and
zkey validation (malformed input is INVALID_ARGUMENT before any device is touched)."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

import icicle_snark_b200 as pkg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _proof_from_json(text):
    d = json.loads(text)
    p = pkg.bindings.Groth16Proof()

    def put(dst, off, val):
        for i, w in enumerate(np.frombuffer(int(val).to_bytes(32, "little"), dtype=np.uint32)):
            dst[off + i] = int(w)

    put(p.pi_a, 0, d["pi_a"][0]); put(p.pi_a, 8, d["pi_a"][1])
    put(p.pi_b, 0, d["pi_b"][0][0]); put(p.pi_b, 8, d["pi_b"][0][1]); put(p.pi_b, 16, d["pi_b"][1][0]); put(p.pi_b, 24, d["pi_b"][1][1])
    put(p.pi_c, 0, d["pi_c"][0]); put(p.pi_c, 8, d["pi_c"][1])
    return p


@pytest.mark.parametrize("name", ["complex_6.proof_r1s1", "complex_6.proof_rs", "complex_100.proof_r1s1", "complex_100.proof_rs"])
def test_c_json_writer_reproduces_golden_bytes(lib, name):
    text = open(os.path.join(GOLD, name + ".json")).read()
    proof = _proof_from_json(text)
    buf = C.create_string_buffer(4096)
    lib.dll.b200_proof_to_json.restype = C.c_size_t
    n = lib.dll.b200_proof_to_json(C.byref(proof), buf, C.c_size_t(4096))
    assert n == len(text) and buf.value.decode() == text
    assert pkg.proof_json(proof) == text  # the Python mirror agrees
    assert lib.dll.b200_proof_to_json(C.byref(proof), buf, C.c_size_t(10)) == 0  # buffer too small
    # zero / small coordinates print as "0" / without leading zeros
    z = pkg.bindings.Groth16Proof()
    z.pi_a[0] = 7
    lib.dll.b200_proof_to_json(C.byref(z), buf, C.c_size_t(4096))
    d = json.loads(buf.value.decode())
    assert d["pi_a"] == ["7", "0", "1"] and d["pi_b"][2] == ["1", "0"] and d["protocol"] == "groth16" and d["curve"] == "bn128"


def test_zkey_validation_happens_before_the_device(lib):
    good = open(os.path.join(GOLD, "complex_6.zkey"), "rb").read()
    handle = C.c_void_p()

    def create(data):
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        return lib.dll.b200_zkey_cache_create(buf, C.c_size_t(len(data)), C.c_int(1), C.byref(handle))

    INVALID_ARGUMENT = 11
    assert create(b"wtns" + good[4:]) == INVALID_ARGUMENT            # wrong magic
    assert create(good[:40]) == INVALID_ARGUMENT                      # truncated
    assert create(good[:4] + struct.pack("<I", 3) + good[8:]) == INVALID_ARGUMENT  # version > 2
    bad_prime = bytearray(good)
    i = good.index(bytes.fromhex("47fd7cd8168c203c"))                  # low limbs of q in the header
    bad_prime[i] ^= 1
    assert create(bytes(bad_prime)) == INVALID_ARGUMENT               # not BN254
    bad_dom = bytearray(good)
    j = good.index(struct.pack("<III", 8, 1, 8))                       # n_vars, n_public, domain_size
    bad_dom[j + 8:j + 12] = struct.pack("<I", 12)                      # not a power of two
    assert create(bytes(bad_dom)) == INVALID_ARGUMENT
    rc = create(good)  # valid file: proceeds to the device (succeeds on a GPU box, INVALID_DEVICE here)
    assert rc in (0, 1)
    if rc == 0:
        lib.dll.b200_zkey_cache_destroy(handle)
    assert lib.dll.b200_zkey_cache_create(None, C.c_size_t(0), C.c_int(1), C.byref(handle)) == 3  # INVALID_POINTER
