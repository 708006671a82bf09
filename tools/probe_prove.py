"""Scratch: build a synthetic instance, run a few proofs (for ncu launch lists)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
n = int(sys.argv[1]); f = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
zkey, wtns, vk = synth.make_complex_circuit(lib, n)
cache = pkg.ZKeyCache(lib, zkey, precompute=f)
nw = cache.n_vars
w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8).copy()
for i in range(reps):
    t = time.time(); p, tm = cache.prove(w, 1, 1); print("prove ms", (time.time() - t) * 1e3, tm.total_ms)
