#!/usr/bin/env python
"""One rank of an N-GPU sharded proof emulated on ONE GPU (timing only: the exchanged slices are whatever the buffer
holds, so the partial sums are not a proof): the rank's own quotient polynomial(s), witness-MSM shard and H shard through
b200_groth16_commit_begin / commit_end.  Used with `ncu --metrics gpu__time_duration.sum` to get the per-kernel list of
a small shard without paying for N GPUs.

  python tools/shard_probe.py --world 8 --rank 0 [--constraints 3200000] [--precompute 16] [--skew 0.049] [--reps 5]
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
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", default="0", help="comma list of ranks to emulate one after the other")
    ap.add_argument("--constraints", type=int, default=3_200_000)
    ap.add_argument("--precompute", type=int, default=16)
    ap.add_argument("--skew", type=float, default=0.049)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    pkg = ge.load_package()
    lib = pkg.lib()
    lib.set_device("CUDA", 0)
    inst = bench.load_instance(args.constraints) or bench.make_instance(lib, args.constraints)
    zkey, wtns, _ = inst
    if args.skew > 0:
        os.environ["B200_SHARD_SKEW"] = repr(args.skew)  # (used by the uniform plan only)
    for rank in (int(x) for x in args.rank.split(",")):
        cache = pkg.ZKeyCache(lib, zkey, precompute=args.precompute, rank=rank, world=args.world)
        nw = cache.n_vars
        w = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8)
        w_dev = torch.from_numpy(w.copy().view(np.int32)).cuda()
        N = cache.domain_size
        first, count = pkg.multi_gpu.owned_polys(rank, args.world)
        lo, hi = cache.h_range()
        mine = torch.zeros((max(count, 1), N, 8), dtype=torch.int32, device="cuda")
        # random field-sized values: h = a.b - c is then a generic scalar vector, as in a real proof
        sl = torch.randint(0, 1 << 28, (3, max(hi - lo, 1), 8), dtype=torch.int32, device="cuda")
        ts = []
        for it in range(args.reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cache.commit_begin(w_dev.data_ptr(), first, count, mine.data_ptr(), n_witness=nw)
            _, tm = cache.commit_end(sl[1].data_ptr(), sl[0].data_ptr(), sl[2].data_ptr())
            torch.cuda.synchronize()
            if it >= 2:
                ts.append((time.perf_counter() - t0) * 1e3)
        print(f"rank {rank}/{args.world} polys {count} ranges {[(hi - lo) for lo, hi in cache.ranges()]}: best {min(ts):.3f} ms median {sorted(ts)[len(ts) // 2]:.3f} ms | "
              f"ntt {tm.ntt_ms:.2f} g1 {tm.msm_g1_ms:.2f} g2 {tm.msm_g2_ms:.2f} total {tm.total_ms:.2f}", flush=True)
        cache.close()
        del mine, sl, w_dev


if __name__ == "__main__":
    main()
