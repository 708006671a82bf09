import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); B = pkg.bindings; lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
rng = np.random.default_rng(5)
sc = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] &= 0x0FFFFFFF
k = rng.integers(0, 1 << 32, size=(1 << 14, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] &= 0x0FFFFFFF
base = synth.fixed_base(lib, k, g2=True)
pts = np.tile(base, (n >> 14, 1))
d_pts = torch.from_numpy(pts.view(np.int32)).cuda(); d_sc = torch.from_numpy(sc.view(np.int32)).cuda()
d_res = torch.zeros(48, dtype=torch.int32, device="cuda")
cfg = B.MSMConfig.default()
cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
cfg.are_points_montgomery_form = True; cfg.is_async = True
for c in [0] + [int(x) for x in sys.argv[2:]]:
    cfg.c = c
    fn = lambda: lib.msm(d_sc.data_ptr(), d_pts.data_ptr(), cfg, g2=True, results=d_res.data_ptr(), msm_size=n)
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"msm g2 2^{lg} c={c}: best {min(ts):.3f} ms  {n/min(ts)/1e3:.1f} Mpts/s")
