"""The reference's own C++ frontend + CPU backend (ICICLE 3.8.0 as vendored in /root/reference/icicle),
compiled by oracle/Makefile.ref into oracle/_ref/libicicle_ref_cpu.so, driven through the SAME ctypes
structs as the product (the ABI is identical - that is the drop-in claim).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs.  Never imported by icicle-snark_b200/.
"""
import os
import subprocess
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libicicle_ref_cpu.so")
REFERENCE_TREE = "/root/reference/icicle"

_ref = None


def build_ref(jobs=8):
    """Only possible where /root/reference is mounted (the build container)."""
    if os.path.exists(REF_LIB):
        return REF_LIB
    if not os.path.isdir(REFERENCE_TREE):
        raise FileNotFoundError(f"{REF_LIB} is not built and {REFERENCE_TREE} is not present on this machine")
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-f", "Makefile.ref", f"-j{jobs}"])
    return REF_LIB


def available():
    return os.path.exists(REF_LIB)


def ref():
    """IcicleLib over the reference CPU library, device set to "CPU"."""
    global _ref
    if _ref is None:
        pkg = sys.modules.get("icicle_snark_b200")
        if pkg is None:
            raise RuntimeError("load the icicle_snark_b200 package first (tests/conftest.py does)")
        class RefLib(pkg.IcicleLib):
            """Tracks the process-wide NTT domain so callers can tell which root is live
            (initialising an existing domain is a no-op in the reference, ntt.cuh:452)."""
            domain_root = None

            def ntt_init_domain(self, root):
                super().ntt_init_domain(root)
                if self.domain_root is None:
                    self.domain_root = bytes(memoryview(root).tobytes())

            def ntt_release_domain(self):
                super().ntt_release_domain()
                self.domain_root = None

        r = RefLib(REF_LIB)
        r.set_device("CPU", 0)
        _ref = r
    return _ref
