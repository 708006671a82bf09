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
    assert declared <= listed | {"b200_launch_count"}, sorted(declared - listed)


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


def test_header_is_plain_c_and_links(tmp_path):
    """include/icicle_b200.h is what a cgo / JNI / Rust-bindgen consumer sees: it must compile as C99 (and C++11) on its
    own, and a C program using it must link against the library and get the documented struct sizes."""
    import subprocess
    src = tmp_path / "consumer.c"
    src.write_text(
        '#include <stdio.h>\n#include "icicle_b200.h"\n'
        "int main(void) {\n"
        "  printf(\"%zu %zu %zu %zu %zu %s\\n\", sizeof(MSMConfig), sizeof(NTTConfig), sizeof(VecOpsConfig),\n"
        "         sizeof(b200_groth16_proof), sizeof(bn254_fq12_t), b200_version());\n"
        "  bn254_scalar_t a = {{5}}, b = {{7}}, c;\n  bn254_mul(&a, &b, &c);\n  return c.limbs[0] == 35 ? 0 : 1;\n}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)
    exe = tmp_path / "consumer"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-licicle_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[:5] == ["40", "64", "32", "256", "384"] and "sm_100a" in out


def test_plain_c_multi_gpu_host_example_compiles_and_links(tmp_path):
    """examples/multi_gpu_host.c: the multi-GPU data plane driven from plain C (fork + pipes for the token), i.e. what a
    Rust / Go / Java host does through its FFI.  Must compile as strict C99 against the public header alone and link against
    the library; without arguments it prints its usage (running it needs GPUs)."""
    import subprocess
    src = os.path.join(ROOT, "examples", "multi_gpu_host.c")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(pkg.LIB_PATH)
    exe = tmp_path / "multi_gpu_host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, src, "-o", str(exe), "-L", libdir,
                    "-licicle_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
