import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); B = pkg.bindings; lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
lg = int(sys.argv[1]); n = 1 << lg
rng = np.random.default_rng(5)
k = rng.integers(0, 1 << 32, size=(1 << 16, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] %= 0x30644e72
base = synth.fixed_base(lib, k)
pts = torch.from_numpy(np.tile(base, (n >> 16, 1)).view(np.int32)).cuda()
res = torch.zeros(24, dtype=torch.int32, device="cuda")
cfg = B.MSMConfig.default(); cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
cfg.are_points_montgomery_form = True; cfg.is_async = True
for name, frac in (("uniform", 0.0), ("50% 0/1", 0.5), ("90% 0/1", 0.9), ("99% 0/1", 0.99), ("all ones", 1.01)):
    sc = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] %= 0x30644e72
    pick = rng.random(n)
    small = pick < frac
    sc[small] = 0
    sc[small, 0] = (rng.random(small.sum()) < 0.5).astype(np.uint32) if frac <= 1 else 1
    d_sc = torch.from_numpy(sc.view(np.int32)).cuda()
    fn = lambda: lib.msm(d_sc.data_ptr(), pts.data_ptr(), cfg, results=res.data_ptr(), msm_size=n)
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"msm g1 2^{lg} {name:10s}: {min(ts):8.3f} ms")
