import ctypes as C, sys
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); lib = pkg.lib(); lib.set_device("CUDA", 0)
lib.dll.b200_pipe_peak.restype = C.c_double
for m, name in enumerate(["imad.lo", "imad.wide", "imad.hi", "dfma", "wide+dfma mixed", "wide carry chain", "Fq mul 8x32 CIOS (G mul/s x1000)", "Fq mul 9x29 imm (G mul/s x1000)", "Fq mul 9x29 regs (G mul/s x1000)", "carry-save wide MAC (+counter)"]):
    print(f"{name:20s} {lib.dll.b200_pipe_peak(m)/1e12:8.3f} Tops/s")
print("mul29 selfcheck mismatches:", lib.dll.b200_mul29_selfcheck())
