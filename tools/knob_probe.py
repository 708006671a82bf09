#!/usr/bin/env python
"""A/B of library environment knobs on the full proof, one process, one instance: for every value of --env NAME=v1,v2,...
a fresh ZKeyCache is built (the library reads its knobs at cache creation / call time) and `--reps` proofs are timed
(host clock around the synchronous call, L2 flushed in between, witness resident in HBM).

  python tools/knob_probe.py --env B200_MSM_QUAD=0 (knobs that are read once per process need one process per value) [--constraints 3200000] [--precompute 16] [--reps 8]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", required=True)
    ap.add_argument("--constraints", type=int, default=3_200_000)
    ap.add_argument("--precompute", type=int, default=16)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    name, vals = args.env.split("=")
    pkg = ge.load_package()
    lib = pkg.lib()
    lib.set_device("CUDA", 0)
    zkey, wtns, _ = bench.load_instance(args.constraints) or bench.make_instance(lib, args.constraints)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ref_json = None
    for v in vals.split(","):
        if v == "unset":
            os.environ.pop(name, None)
        else:
            os.environ[name] = v
        t0 = time.time()
        cache = pkg.ZKeyCache(lib, zkey, precompute=args.precompute)
        lib.device_synchronize()
        t_build = time.time() - t0
        nw = cache.n_vars
        w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8)
        w_dev = torch.from_numpy(w.copy().view(np.int32)).cuda()
        ts = []
        for it in range(args.reps + 3):
            flush.fill_(it & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proof, tm = cache.prove(w_dev.data_ptr(), bench.R_BLIND, bench.S_BLIND, n_witness=nw)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append((time.perf_counter() - t0) * 1e3)
        js = pkg.proof_json(proof)
        ref_json = ref_json or js
        print(f"{name}={v}: median {sorted(ts)[len(ts) // 2]:.3f} ms best {min(ts):.3f} | ntt {tm.ntt_ms:.2f} g1 {tm.msm_g1_ms:.2f} "
              f"g2 {tm.msm_g2_ms:.2f} total {tm.total_ms:.2f} | cache build {t_build:.2f} s | same proof: {js == ref_json}", flush=True)
        cache.close()
        del w_dev


if __name__ == "__main__":
    main()
