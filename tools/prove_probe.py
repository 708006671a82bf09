#!/usr/bin/env python
"""A few warm proofs of the bench instance and nothing else - the short command to wrap in `ncu --set full -k regex:<kernel>`.

  python tools/prove_probe.py [--constraints 3200000] [--precompute 16] [--reps 3]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--constraints", type=int, default=3_200_000)
    ap.add_argument("--precompute", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    pkg = ge.load_package()
    lib = pkg.lib()
    lib.set_device("CUDA", 0)
    zkey, wtns, _ = bench.load_instance(args.constraints) or bench.make_instance(lib, args.constraints)
    cache = pkg.ZKeyCache(lib, zkey, precompute=args.precompute)
    nw = cache.n_vars
    w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8)
    w_dev = torch.from_numpy(w.copy().view(np.int32)).cuda()
    for _ in range(args.reps):
        _, tm = cache.prove(w_dev.data_ptr(), bench.R_BLIND, bench.S_BLIND, n_witness=nw)
        torch.cuda.synchronize()
    print(f"total {tm.total_ms:.2f} ms", flush=True)
    cache.close()


if __name__ == "__main__":
    main()
