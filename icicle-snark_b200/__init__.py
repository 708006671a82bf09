"""icicle-snark_b200: B200-native Groth16 proving hot path behind icicle-snark's API.

Python host-side mirror of the reference's Rust layers for the hot path (the container has no
Rust toolchain; see INTEGRATION.md for the Rust-side binding):

  bindings.IcicleLib   <-> wrappers/rust/icicle-{runtime,core,curves/icicle-bn254}  (op-level C ABI)
  prover.groth16_prove <-> src/lib.rs:33-61, src/cache.rs CacheManager               (fused C ABI)
  prover.groth16_verify <-> src/lib.rs:63-82, src/proof_helper.rs:319-372            (host pairing, as in the reference)

All arithmetic is in csrc/ (hand-written sm_100a CUDA behind lib/libicicle_b200.so).  There is no
CPU fallback: loading fails loudly if the library has not been built.

The directory name contains a hyphen (it mirrors the reference's repo name), so import it through
`load_package()` in __graft_entry__.py / tests/conftest.py, which registers it as `icicle_snark_b200`.
"""
import os
import subprocess

from .bindings import *  # noqa: F401,F403
from .bindings import IcicleLib

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libicicle_b200.so")

_lib = None


def build(jobs=8, verbose=False):
    """Compile csrc/*.cu for sm_100a into lib/libicicle_b200.so (nvcc; no GPU needed)."""
    r = subprocess.run(["make", "-C", PKG_DIR, f"-j{jobs}"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libicicle_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def lib() -> IcicleLib:
    """The product library (process-wide singleton)."""
    global _lib
    if _lib is None:
        _lib = IcicleLib(LIB_PATH)
    return _lib

_tools = None


def tools_lib():
    """lib/libicicle_b200_tools.so: pipe probes, multiplier microbenchmarks and host models (research/); used by
    bench.py for the measured integer-pipe peak and by the tests of the models - never by the product path."""
    global _tools
    if _tools is None:
        import ctypes as C
        path = os.path.join(PKG_DIR, "lib", "libicicle_b200_tools.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: build it first (make -C {PKG_DIR})")
        _tools = C.CDLL(path)
        for name in ("b200_probe_cycles", "b200_imad_wide_peak", "b200_pipe_peak"):
            getattr(_tools, name).restype = C.c_double
        _tools.b200_probe_name.restype = C.c_char_p
    return _tools


from .prover import (CacheManager, ZKeyCache, groth16_prove, groth16_verify, groth16_verify_points,  # noqa: E402,F401
                     proof_json, proof_to_dict)
from . import multi_gpu  # noqa: E402,F401
