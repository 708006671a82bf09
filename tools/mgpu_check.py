#!/usr/bin/env python
"""Multi-GPU parity check of the in-library data plane (b200_comm_*, b200_groth16_prove_sharded, b200_msm_sharded),
run under torchrun with one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py

1. the committed golden instances (tests/golden/complex_{6,100}) proved over N GPUs with fixed (r, s): rank 0's
   proof.json must equal the golden file byte for byte (the golden files come from the reference pipeline,
   tests/golden/make_golden.py), with and without precompute tables, witness in host and in device memory;
2. a freshly generated instance (default 20 000 constraints): sharded proof == the single-GPU proof of the full cache;
3. b200_msm_sharded (G1 and G2, 2^14 + 3 points): the total must equal bn254_msm over all points on one GPU.
Prints "MGPU_CHECK_OK world=N" on rank 0 and exits 0 on every rank when everything matches.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
B = pkg.bindings
from tools import synth  # noqa: E402

# the blinding factors the golden proof_rs.json files were made with (tests/test_groth16_oracle.py)
FIXED_R = 0x1d2c3b4a5968778695a4b3c2d1e0f00112233445566778899aabbccddeeff001 % synth.R
FIXED_S = 0x0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcdef % synth.R


def wtns_words(raw, n):
    return np.frombuffer(raw, dtype=np.uint32, count=n * 8, offset=len(raw) - n * 32).reshape(n, 8).copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fresh", type=int, default=20000, help="constraints of the freshly generated instance (0 = skip)")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.lib()
    lib.set_device("CUDA", local)
    comm = pkg.multi_gpu.LibComm.from_torch(lib)
    assert (comm.rank, comm.world) == (rank, world)
    R, S = FIXED_R, FIXED_S
    fails = []

    # 1. golden instances
    for n in (6, 100):
        base = os.path.join(ROOT, "tests", "golden", f"complex_{n}")
        zkey, wtns = open(base + ".zkey", "rb").read(), open(base + ".wtns", "rb").read()
        gold = open(base + ".proof_rs.json").read()
        for precompute in (1, 16):
            cache = pkg.ZKeyCache(lib, zkey, precompute=precompute, rank=rank, world=world)
            w = wtns_words(wtns, cache.n_vars)
            w_dev = torch.from_numpy(w.view(np.int32)).cuda()
            for src, ptr in (("host", w), ("device", w_dev.data_ptr())):
                for it in range(2):  # second proof on the same cache: buffers reused
                    proof, _ = cache.prove_sharded(comm, ptr, R, S, n_witness=cache.n_vars)
                    if rank == 0 and pkg.proof_json(proof) != gold:
                        fails.append(f"golden complex_{n} precompute={precompute} witness={src} iteration {it}")
            cache.close()

    # 2. fresh instance: sharded == single GPU
    if args.fresh:
        zkey, wtns, _ = synth.make_complex_circuit(lib, args.fresh)  # deterministic: every rank builds the same files
        cache = pkg.ZKeyCache(lib, zkey, precompute=4, rank=rank, world=world)
        w = wtns_words(wtns, cache.n_vars)
        proof, _ = cache.prove_sharded(comm, w, R, S, n_witness=cache.n_vars)
        cache.close()
        if rank == 0:
            full = pkg.ZKeyCache(lib, zkey, precompute=4)
            want, _ = full.prove(w, R, S)
            full.close()
            if pkg.proof_json(proof) != pkg.proof_json(want):
                fails.append(f"fresh instance ({args.fresh} constraints): sharded proof differs from the single-GPU proof")

    # 3. sharded standalone MSM
    rng = np.random.default_rng(99)
    n = (1 << 14) + 3
    sc = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    sc[:, 7] %= 0x30644e72
    for g2 in (False, True):
        pts = lib.generate_affine_points(n, g2=g2)
        t = torch.from_numpy(pts.view(np.int32).copy()).cuda()
        dist.broadcast(t, src=0)  # every rank the same points (the generator is random)
        pts = t.cpu().numpy().view(np.uint32)
        lo, hi = pkg.multi_gpu.shard_range(n, rank, world)
        cfg = B.MSMConfig.default()
        got = comm.msm(np.ascontiguousarray(sc[lo:hi]), np.ascontiguousarray(pts[lo:hi]), hi - lo, cfg, g2=g2)
        want = lib.msm(sc, pts, g2=g2)[0]
        if not np.array_equal(lib.to_affine(got, g2=g2), lib.to_affine(want, g2=g2)):
            fails.append(f"b200_msm_sharded g2={g2} on rank {rank}")

    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    for f in fails:
        print(f"[rank {rank}] MISMATCH: {f}", flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print(f"MGPU_CHECK_OK world={world}", flush=True)


if __name__ == "__main__":
    main()
