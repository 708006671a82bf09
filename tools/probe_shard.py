"""Scratch: one rank of an N-way sharded proof emulated on ONE GPU (rank 0 of `world`), swept over the MSM window
width (B200_MSM_C).  Prints the device time of commit_partials per configuration."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); lib = pkg.lib(); lib.set_device("CUDA", 0)
from tools import synth
n = int(sys.argv[1]); worlds = [int(x) for x in sys.argv[2].split(",")]; cs = [int(x) for x in sys.argv[3].split(",")]
zkey, wtns, vk = synth.make_complex_circuit(lib, n)
nw = None
for world in worlds:
    for c in cs:
        if c:
            os.environ["B200_MSM_C"] = str(c)
        else:
            os.environ.pop("B200_MSM_C", None)
        cache = pkg.ZKeyCache(lib, zkey, precompute=16, rank=0, world=world)
        nw = cache.n_vars
        w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8).copy()
        wd = torch.from_numpy(w.view(np.int32)).cuda()  # device-resident witness: the timing excludes the host copy
        ts = []
        for i in range(4):
            parts, tm = cache.commit_partials(wd.data_ptr(), n_witness=nw)
            ts.append((tm.total_ms, tm.msm_g1_ms, tm.msm_g2_ms, tm.ntt_ms))
        best = min(ts)
        print(f"world {world} c {c or 'auto'}: total {best[0]:.2f} ms  g1 {best[1]:.2f} g2 {best[2]:.2f} ntt {best[3]:.2f}", flush=True)
        cache.close()
