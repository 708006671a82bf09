#!/usr/bin/env python
"""Standalone sweeps (BASELINE.json configs[4]): BN254 G1/G2 MSM 2^16..2^26 points and Fr NTT 2^16..2^24,
inputs resident in HBM, CUDA-event timing, one JSON line per point.  Under torchrun (N ranks) the MSM is
sharded by contiguous point ranges through the library's own entry point b200_msm_sharded (each rank runs the full
Pippenger on n/N points; one 96 B/192 B ncclAllGather inside the library + N-1 host adds on every rank) and the NTT
runs as N replicas (it does not shard).

  python tools/sweep.py [--msm 16,18,...] [--g2 16,18] [--ntt 16,...] [--precompute F]
  python -m torch.distributed.run --nproc-per-node N tools/sweep.py ...
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
B = pkg.bindings
from tools import synth  # noqa: E402


def timed(fn, reps=5, warm=2, sync=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if sync:
            sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def run(lib, pkg=pkg, msm=(16, 18, 20, 22, 24, 26), g2=(16, 18, 20, 22), ntt=(16, 18, 20, 22, 24), precompute=1, emit=None, log=None):
    """Runs the sweeps on this rank's device; returns the list of result dicts (rank 0) and calls emit(dict) per point.
    Under torch.distributed (initialised by the caller) the MSM is sharded, the NTT replicated."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    rng = np.random.default_rng(20261017)
    results = []
    comm = pkg.multi_gpu.LibComm.from_torch(lib) if world > 1 else None

    def out(d):
        if rank == 0:
            results.append(d)
            if emit:
                emit(d)

    def scalars(n):
        s = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        s[:, 7] %= 0x30644e72  # uniform below r up to the top limb
        return s

    def run_msm(lg, is_g2):
        n = 1 << lg
        n_loc = n // world
        # points: k_i * G from the fixed-base tool (distinct, valid), Montgomery form, tiled above 2^22
        m = min(n_loc, 1 << 22 if not is_g2 else 1 << 20)
        base = synth.fixed_base(lib, scalars(m), g2=is_g2)
        reps = (n_loc + m - 1) // m
        pts = torch.from_numpy(base.view(np.int32)).cuda()
        if reps > 1:
            pts = pts.repeat(reps, 1)[:n_loc].contiguous()
        sc = torch.from_numpy(scalars(n_loc).view(np.int32)).cuda()
        res = torch.zeros(48 if is_g2 else 24, dtype=torch.int32, device="cuda")
        cfg = B.MSMConfig.default()
        cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
        cfg.are_points_montgomery_form = True
        cfg.is_async = True
        tab = pts
        if precompute > 1:
            cfg.precompute_factor = precompute
            w = 32 if is_g2 else 16
            tab = torch.empty((n_loc * precompute, w), dtype=torch.int32, device="cuda")
            lib.msm_precompute_bases(pts.data_ptr(), cfg, g2=is_g2, n=n_loc, out=tab.data_ptr())
            torch.cuda.synchronize()
        def step():
            if comm is not None:  # synchronous: partial sum, all-gather and fold inside the library
                comm.msm(sc.data_ptr(), tab.data_ptr(), n_loc, cfg, g2=is_g2)
            else:
                lib.msm(sc.data_ptr(), tab.data_ptr(), cfg, g2=is_g2, results=res.data_ptr(), msm_size=n_loc)

        ms = timed(step, sync=(dist.barrier if world > 1 else None))
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        info = (C.c_int32 * 8)()
        lib.dll.b200_msm_plan_info(C.c_int(n_loc), C.c_int(0), C.c_int(254), C.c_int(precompute), C.c_int(int(is_g2)), info, None)
        out({"op": "msm_g2" if is_g2 else "msm_g1", "log_n": lg, "n_gpus": world, "precompute": precompute, "c": int(info[0]),
             "windows": int(info[1]), "ms": round(ms, 4), "mpoints_s": round(n / ms / 1e3, 2),
             # wide multiply-adds per second over the bucket additions alone (SURVEY 8d unit: 10 / 30 products x 136 MACs)
             "t_mac_s": round(n * int(info[1]) * (30 if is_g2 else 10) * 136 / ms / 1e9, 3)})
        del pts, sc, tab

    for lg in msm:
        if log:
            log(f"sweep: G1 MSM 2^{lg}")
        run_msm(lg, False)
    for lg in g2:
        if log:
            log(f"sweep: G2 MSM 2^{lg}")
        run_msm(lg, True)
    ntt = list(ntt)
    if ntt:
        lib.ntt_release_domain()
        lib.ntt_init_domain(lib.get_root_of_unity(1 << max(ntt)))
    for lg in ntt:
        n, batch = 1 << lg, 3
        x = torch.from_numpy(scalars(n * batch).view(np.int32)).cuda()
        y = torch.empty_like(x)
        cfg = B.NTTConfig.default()
        cfg.batch_size = batch
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = True
        cfg.is_async = True
        ms = timed(lambda: lib.ntt(x.data_ptr(), B.kForward, cfg, out=y.data_ptr(), size=n))
        muls = batch * n * (lg / 2 + (-(-lg // 8) - 1))
        out({"op": "ntt_fr_fwd_batch3", "log_n": lg, "n_gpus": world, "replicas": world, "ms": round(ms, 4),
             "melem_s": round(batch * n / ms / 1e3, 1), "algorithmic_gb_s": round(batch * n * 64 / ms / 1e6, 1),
             "field_mul_g_s": round(muls / ms / 1e6, 1), "t_mac_s": round(muls * 136 / ms / 1e9, 3)})
        del x, y
    if comm is not None:
        comm.close()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msm", default="16,18,20,22,24,26")
    ap.add_argument("--g2", default="16,18,20,22")
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--precompute", type=int, default=1)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.lib()
    lib.set_device("CUDA", local)
    ints = lambda s: [int(x) for x in s.split(",") if x]
    run(lib, pkg, ints(args.msm), ints(args.g2), ints(args.ntt), args.precompute, emit=lambda d: print(json.dumps(d), flush=True))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
