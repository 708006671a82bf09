"""Scratch perf probe (not the bench): device-resident MSM / NTT timings with CUDA events."""
import ctypes as C
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
B = pkg.bindings
lib = pkg.lib()
lib.set_device("CUDA", 0)
lib.dll.b200_imad_peak.restype = C.c_double
print("imad lo  Gops/s", lib.dll.b200_imad_peak(0) / 1e9)
print("imad wide Gops/s", lib.dll.b200_imad_peak(1) / 1e9)

def dev(arr):
    t = torch.from_numpy(arr.view(np.int32)).cuda()
    return t

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

rng = np.random.default_rng(5)
logs = [int(a) for a in sys.argv[1:]] or [16, 20, 22]
nmax = 1 << max(logs)
t0 = time.time()
pts = lib.generate_affine_points(nmax)
print("gen points s", time.time() - t0)
sc = rng.integers(0, 1 << 32, size=(nmax, 8), dtype=np.uint64).astype(np.uint32)
sc[:, 7] &= 0x0FFFFFFF
d_pts_std = dev(pts)
d_pts = torch.empty_like(d_pts_std)
cfgv = B.VecOpsConfig.default(); cfgv.is_a_on_device = cfgv.is_result_on_device = True
B.check(lib.dll.bn254_affine_convert_montgomery(C.c_void_p(d_pts_std.data_ptr()), C.c_size_t(nmax), C.c_bool(True), C.byref(cfgv), C.c_void_p(d_pts.data_ptr())))
d_sc = dev(sc)
d_res = torch.zeros(24, dtype=torch.int32, device="cuda")
for lg in logs:
    n = 1 << lg
    for c in ([0] if len(sys.argv) < 99 else [0]):
        cfg = B.MSMConfig.default()
        cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
        cfg.are_points_montgomery_form = True
        cfg.is_async = True
        cfg.c = c
        fn = lambda: lib.msm(d_sc.data_ptr(), d_pts.data_ptr(), cfg, results=d_res.data_ptr(), msm_size=n)
        best, med = timeit(fn)
        print(f"msm g1 2^{lg} c={c}: best {best:.3f} ms med {med:.3f} ms  {n/best/1e3:.1f} Mpts/s")
# NTT
lib.ntt_init_domain(lib.get_root_of_unity(1 << 24))
for lg in [16, 20, 22, 24]:
    n = 1 << lg
    batch = 3
    x = torch.randint(0, 1 << 28, (n * batch, 8), dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    cfg = B.NTTConfig.default(); cfg.batch_size = batch
    cfg.are_inputs_on_device = cfg.are_outputs_on_device = True; cfg.is_async = True
    fn = lambda: lib.ntt(x.data_ptr(), B.kForward, cfg, out=y.data_ptr(), size=n)
    best, med = timeit(fn)
    print(f"ntt 2^{lg} x{batch}: best {best:.3f} ms  ({batch*n*64/best/1e6:.0f} GB/s algorithmic)")
    fn = lambda: lib.ntt(x.data_ptr(), B.kInverse, cfg, out=x.data_ptr(), size=n)
    best, med = timeit(fn)
    print(f"intt inplace 2^{lg} x{batch}: best {best:.3f} ms")
# vec mul
n = 1 << 22
a = torch.randint(0, 1 << 28, (n, 8), dtype=torch.int32, device="cuda"); b = a.clone(); o = torch.empty_like(a)
cfgv = B.VecOpsConfig.default(); cfgv.is_a_on_device = cfgv.is_b_on_device = cfgv.is_result_on_device = True; cfgv.is_async = True
fn = lambda: B.check(lib.dll.bn254_vector_mul(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_uint64(n), C.byref(cfgv), C.c_void_p(o.data_ptr())))
best, med = timeit(fn)
print(f"vec mul 2^22: {best:.3f} ms {n*96/best/1e6:.0f} GB/s")
