"""Scratch: one rank's share of a sharded proof on a single GPU (world emulated)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
n = int(sys.argv[1]); f = int(sys.argv[2]); world = int(sys.argv[3])
zkey, wtns, vk = synth.make_complex_circuit(lib, n)
cache = pkg.ZKeyCache(lib, zkey, precompute=f, rank=0, world=world)
nw = cache.n_vars
w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8).copy()
for i in range(3):
    t = time.time(); p, tm = cache.commit_partials(w); print("commit ms", (time.time() - t) * 1e3, tm.total_ms, tm.msm_g1_ms, tm.msm_g2_ms, tm.ntt_ms)
