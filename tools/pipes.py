import ctypes as C, sys
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); lib = pkg.lib(); lib.set_device("CUDA", 0)
lib.dll.b200_pipe_peak.restype = C.c_double
for m, name in enumerate(["imad.lo", "imad.wide", "imad.hi", "dfma", "wide+dfma mixed", "wide carry chain", "Fq mul 8x32 CIOS (G mul/s x1000)", "Fq mul 9x29 imm (G mul/s x1000)", "Fq mul 9x29 regs (G mul/s x1000)", "carry-save wide MAC (+counter)", "Fq mul FP64 6x48 alone (G mul/s x1000)", "CIOS + FP64 co-run, half the warps each (G mul/s x1000)"]):
    print(f"{name:20s} {lib.dll.b200_pipe_peak(m)/1e12:8.3f} Tops/s")
print("mul29 selfcheck mismatches:", lib.dll.b200_mul29_selfcheck())
print("mul48 (FP64) selfcheck mismatches:", lib.dll.b200_mul48_selfcheck())
for m, name in ((12, "Fq sqr dedicated (SOS)"), (13, "Fq mul wide + SOS reduce")):
    print(f"{name:40s} {lib.dll.b200_pipe_peak(m)/1e9:8.1f} G/s")
for ci, cf in ((4, 2),):
    out = (C.c_double * 3)()
    lib.dll.b200_corun_test(out, ci, cf)
    print(f"co-residency {ci} int CTAs + {cf} fp64 CTAs per SM: int alone {out[0]/1e9:.1f}  fp64 alone {out[1]/1e9:.1f}  together {out[2]/1e9:.1f} G mul/s")
