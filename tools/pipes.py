"""Integer-pipe probes of the GPU this runs on (research/pipes2.cu in lib/libicicle_b200_tools.so): cycles per
warp-instruction per SM sub-partition for the instruction classes the field multiplier is built from, the resulting
wide-MAC peak (the MSM/NTT roofline denominator) and the throughput of the production Montgomery product."""
import sys
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
pkg.lib().set_device("CUDA", 0)
t = pkg.tools_lib()
for k in range(t.b200_probe_count()):
    i = t.b200_probe_id(k)
    print(f"T{i:<3d} {t.b200_probe_name(k).decode():58s} {t.b200_probe_cycles(i):7.3f} cycles / warp-instruction / SMSP")
print(f"IMAD.WIDE peak {t.b200_imad_wide_peak() / 1e12:.2f} T wide MAC/s")
print(f"Fq Montgomery products (8x32 CIOS, dependent chains): {t.b200_pipe_peak(6) / 1e9:.1f} G/s")
