import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); B = pkg.bindings; lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
lg = int(sys.argv[1]); f = int(sys.argv[2]); c = int(sys.argv[3]); g2 = len(sys.argv) > 4
n = lg if lg > 64 else 1 << lg
rng = np.random.default_rng(5)
sc = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] &= 0x0FFFFFFF
k = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] &= 0x0FFFFFFF
pts = synth.fixed_base(lib, k, g2=g2)
W = 32 if g2 else 16
d_pts = torch.from_numpy(pts.view(np.int32)).cuda(); d_sc = torch.from_numpy(sc.view(np.int32)).cuda()
d_res = torch.zeros(48, dtype=torch.int32, device="cuda")
cfg = B.MSMConfig.default()
cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
cfg.are_points_montgomery_form = True; cfg.is_async = True
cfg.precompute_factor = f; cfg.c = c
if f > 1:
    d_tab = torch.empty((n * f, W), dtype=torch.int32, device="cuda")
    t0 = time.time()
    lib.msm_precompute_bases(d_pts.data_ptr(), cfg, g2=g2, n=n, out=d_tab.data_ptr()); torch.cuda.synchronize()
    print("precompute s", time.time() - t0)
else:
    d_tab = d_pts
fn = lambda: lib.msm(d_sc.data_ptr(), d_tab.data_ptr(), cfg, g2=g2, results=d_res.data_ptr(), msm_size=n)
for _ in range(2): fn()
torch.cuda.synchronize(); ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"msm {'g2' if g2 else 'g1'} 2^{lg} f={f} c={c}: best {min(ts):.3f} ms  {n/min(ts)/1e3:.1f} Mpts/s")
