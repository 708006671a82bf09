"""The C-ABI library loads and exports every symbol include/icicle_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import icicle_snark_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "icicle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b((?:bn254|icicle|b200)_\w+|\w*config_extension\w*)\s*\(", src))
    return sorted(names)


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) > 90
    missing = [n for n in names if not lib.has(n)]
    assert not missing, f"declared in include/icicle_b200.h but not exported: {missing}"


def test_binding_lists_match_header(lib):
    declared = set(declared_symbols())
    listed = set(pkg.bindings.ABI_SYMBOLS) | set(pkg.bindings.B200_SYMBOLS)
    assert declared <= listed | {"b200_imad_peak", "b200_launch_count"}, sorted(declared - listed)


def test_reference_library_exports_the_same_op_level_abi(ref):
    # drop-in: every op-level symbol we export exists in the reference's own libraries too
    missing = [n for n in pkg.bindings.ABI_SYMBOLS if not ref.has(n)]
    assert not missing, missing


def test_struct_layouts():
    b = pkg.bindings
    assert C.sizeof(b.Device) == 68 and b.Device.id.offset == 64
    assert C.sizeof(b.MSMConfig) == 40 and b.MSMConfig.is_async.offset == 30
    assert C.sizeof(b.NTTConfig) == 64 and b.NTTConfig.coset_gen.offset == 8 and b.NTTConfig.is_async.offset == 54
    assert C.sizeof(b.VecOpsConfig) == 32 and b.VecOpsConfig.columns_batch.offset == 16
    assert C.sizeof(b.Groth16Proof) == 64 + 128 + 64 and C.sizeof(b.Groth16Partials) == 4 * 96 + 192


def test_version_string(lib):
    lib.dll.b200_version.restype = C.c_char_p
    assert b"sm_100a" in lib.dll.b200_version()


def test_no_cpu_backend_behind_the_product(lib):
    # device "CPU" must be refused: no CPU fallback (north_star)
    d = pkg.bindings.Device.new("CPU", 0)
    assert lib.dll.icicle_set_device(C.byref(d)) == 1  # INVALID_DEVICE
