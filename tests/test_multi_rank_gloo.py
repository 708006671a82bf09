"""N > 1 host logic on CPU: two gloo ranks each commit to their contiguous shard (here with the reference
CPU library standing in for the GPU kernels), all_gather the 576 B partials, rank 0 folds them with the
product library's host helpers; the fold equals the unsharded commitments."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest  # noqa: F401
    import icicle_snark_b200 as pkg
    import torch.distributed as dist
    from oracle import ref_cpu
    from util import rand_scalars
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref, lib = ref_cpu.ref(), pkg.lib()
        rng = np.random.default_rng(99)  # same inputs on every rank
        n = 257
        sc, _ = rand_scalars(rng, n)
        p1 = ref.generate_affine_points(4)[np.arange(n) % 4].copy()
        p1 = np.ascontiguousarray(ref.convert_montgomery(ref.convert_montgomery(p1, True, kind="affine"), False, kind="affine"))
        # deterministic across ranks: derive points from the scalars instead of the library's RNG
        gen = ref.generator()
        base = [ref.to_affine(ref.mul_scalar(gen, sc[i])) for i in range(8)]
        p1 = np.array([base[i % 8] for i in range(n)], dtype=np.uint32)
        g2gen = ref.generator(g2=True)
        base2 = [ref.to_affine(ref.mul_scalar(g2gen, sc[i], g2=True), g2=True) for i in range(4)]
        p2 = np.array([base2[i % 4] for i in range(n)], dtype=np.uint32)
        lo, hi = pkg.multi_gpu.shard_range(n, rank, world)
        parts = pkg.bindings.Groth16Partials()
        g1 = ref.msm(np.ascontiguousarray(sc[lo:hi]), np.ascontiguousarray(p1[lo:hi]))[0]
        g2 = ref.msm(np.ascontiguousarray(sc[lo:hi]), np.ascontiguousarray(p2[lo:hi]), g2=True)[0]
        for name in ("a", "b1", "c", "h"):
            getattr(parts, name)[:] = [int(x) for x in g1]
        parts.b2[:] = [int(x) for x in g2]
        allp = pkg.multi_gpu.all_gather_partials(parts, "cpu")
        assert len(allp) == world
        if rank == 0:
            acc1 = np.array(list(allp[0].a), dtype=np.uint32)
            acc2 = np.array(list(allp[0].b2), dtype=np.uint32)
            for p in allp[1:]:
                acc1 = lib.ecadd(acc1, np.array(list(p.a), dtype=np.uint32))
                acc2 = lib.ecadd(acc2, np.array(list(p.b2), dtype=np.uint32), g2=True)
            full1, full2 = ref.msm(sc, p1)[0], ref.msm(sc, p2, g2=True)[0]
            ok = bool(np.array_equal(ref.to_affine(acc1), ref.to_affine(full1)) and
                      np.array_equal(ref.to_affine(acc2, g2=True), ref.to_affine(full2, g2=True)))
            q.put(ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_fold():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_partition():
    import icicle_snark_b200 as pkg
    for n in (0, 1, 7, 3_200_002):
        for world in (1, 2, 3, 8):
            r = [pkg.multi_gpu.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
